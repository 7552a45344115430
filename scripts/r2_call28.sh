#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c28_tests.txt 2>&1; tail -3 gpurun_out/r2c28_tests.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c28_bench_ref.json 2> gpurun_out/r2c28_bench_ref.err
python - <<'PY'
import json
a=json.load(open('gpurun_out/r2c28_bench.json')); b=json.load(open('gpurun_out/r2c28_bench_ref.json'))
print(a['value'], a['ms_per_step'], a['e2e']['value'], '| ref', b['value'], '| same config:', a['config']==b['config'])
print({k:(v.get('value'),v.get('ms_per_step')) for k,v in a['workloads'].items()})
PY
