#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c16_tests.txt 2>&1; tail -4 gpurun_out/r2c16_tests.txt
run() { name=$1; shift; for cfg in "gum 256" "s50 512"; do echo "== $name $cfg: $(env "$@" timeout 120 python scripts/profile_step.py 0 $cfg quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done; }
run A_default PPB_X=0
run B_nofusepack PPB_FUSE_PACK=0
run A2_default PPB_X=0
timeout 400 python bench.py > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c16_bench.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_stats_rank0'],d['e2e']['value']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
