#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c18_tests.txt 2>&1; tail -3 gpurun_out/r2c18_tests.txt
timeout 400 python bench.py > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c18_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
