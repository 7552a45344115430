#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c26_tests.txt 2>&1; tail -4 gpurun_out/r2c26_tests.txt | cut -c1-200
PPB_PERSISTENT=0 timeout 100 python scripts/gemm_sat.py 2>&1 | tail -1
PPB_PERSISTENT=1 timeout 100 python scripts/gemm_sat.py 2>&1 | tail -1
for v in 0 1; do echo "== persistent=$v s50: $(PPB_PERSISTENT=$v timeout 120 python scripts/profile_step.py 0 s50 512 quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done
timeout 400 python bench.py > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c26_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()});print(d['extra']['gate_gemm_saturating_4096x2048x512'])"
