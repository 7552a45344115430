#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c21_tests.txt 2>&1; tail -8 gpurun_out/r2c21_tests.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c21_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
