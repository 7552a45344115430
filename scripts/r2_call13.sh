#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PPB_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c13_tests_pdl.txt 2>&1; tail -5 gpurun_out/r2c13_tests_pdl.txt
for v in 0 1; do
  PPB_PDL=$v timeout 300 python bench.py --no-extra > gpurun_out/r2c13_bench_pdl$v.json 2> gpurun_out/r2c13_bench_pdl$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r2c13_bench_pdl$v.json'));print('PDL=$v', d['value'],d['ms_per_step'],d['ms_per_step_stats_rank0'],d['e2e']['value']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
done
PPB_PDL=1 timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c13_prof_gum_pdl.txt 2>&1; grep "tc launch\|^==\|graph replay\|avg step" gpurun_out/r2c13_prof_gum_pdl.txt | cut -c1-275
PPB_PDL=1 timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c13_prof_s50_pdl.txt 2>&1; head -9 gpurun_out/r2c13_prof_s50_pdl.txt | cut -c1-150; grep "tc launch\|^==\|graph replay" gpurun_out/r2c13_prof_s50_pdl.txt | awk '/^==/{n=0} {n++} n<=9 || (n>=61 && n<=64)' | cut -c1-275
