#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c9_tests.txt 2>&1; tail -3 gpurun_out/r2c9_tests.txt
timeout 400 python bench.py > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c9_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print({k:v.get('value') for k,v in d['workloads'].items()})"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c9_bench_ref.json 2> gpurun_out/r2c9_bench_ref.err; cut -c1-300 gpurun_out/r2c9_bench_ref.json
bash scripts/r2_call8.sh
