#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c11_prof_s50.txt 2>&1; head -12 gpurun_out/r2c11_prof_s50.txt | cut -c1-150; grep "tc launch" gpurun_out/r2c11_prof_s50.txt | sed -n '1,9p;60,64p' | cut -c1-260
timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c11_prof_gum.txt 2>&1; head -24 gpurun_out/r2c11_prof_gum.txt | cut -c1-150; grep "tc launch" gpurun_out/r2c11_prof_gum.txt | cut -c1-260
timeout 300 python scripts/ab_optin.py mix > gpurun_out/r2c11_ab_mix.json 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c11_ab_mix.json'))['mixture_scoring']
for k,v in d.items(): print(k, {a:(round(b['ms'],4), round(b['gbs'])) if isinstance(b,dict) and 'ms' in b else b for a,b in v.items()})
PY
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_model_gpu.py::test_marsaglia_inference_compilation > gpurun_out/r2c11_tests.txt 2>&1; tail -5 gpurun_out/r2c11_tests.txt
timeout 300 python bench.py > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c11_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print({k:v.get('value') for k,v in d['workloads'].items()});print([(x['kernel'],round(x['frac_of_hbm'],3)) for x in d['extra']['scoring_hbm_roofline']])"
