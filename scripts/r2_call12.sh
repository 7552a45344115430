#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c12_tests.txt 2>&1; tail -5 gpurun_out/r2c12_tests.txt
timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c12_prof_s50.txt 2>&1; head -10 gpurun_out/r2c12_prof_s50.txt | cut -c1-150; grep "tc launch\|^==\|graph replay" gpurun_out/r2c12_prof_s50.txt | awk '/^==/{n=0} {n++} n<=9 || (n>=61 && n<=64)' | cut -c1-275
timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c12_prof_gum.txt 2>&1; head -24 gpurun_out/r2c12_prof_gum.txt | cut -c1-150; grep "tc launch\|^==\|graph replay" gpurun_out/r2c12_prof_gum.txt | cut -c1-275
timeout 400 python bench.py > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c12_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print({k:(v.get('value'),v.get('ms_per_call')) for k,v in d['workloads'].items()});print([(x['kernel'],round(x['frac_of_hbm'],3)) for x in d['extra']['scoring_hbm_roofline']])"
