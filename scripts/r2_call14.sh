#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c14_tests.txt 2>&1; tail -5 gpurun_out/r2c14_tests.txt
timeout 400 python bench.py > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c14_bench.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_stats_rank0'],d['e2e']['value']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c14_prof_s50.txt 2>&1; head -12 gpurun_out/r2c14_prof_s50.txt | cut -c1-150; grep "tc launch\|^==\|graph replay" gpurun_out/r2c14_prof_s50.txt | awk '/^==/{n=0} {n++} n<=8 || (n>=61 && n<=63)' | cut -c1-275
timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c14_prof_gum.txt 2>&1; grep "avg step\|graph replay" gpurun_out/r2c14_prof_gum.txt
