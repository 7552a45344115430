#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_cell or baseline or pdl or network" > gpurun_out/r2c24_tests.txt 2>&1; tail -3 gpurun_out/r2c24_tests.txt | cut -c1-200
PPB_PDL=0 timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c24_cupti_s50_nopdl.txt 2>&1; head -8 gpurun_out/r2c24_cupti_s50_nopdl.txt | cut -c1-150; grep "tc launch" gpurun_out/r2c24_cupti_s50_nopdl.txt | sed -n '6,8p;70,72p' | cut -c1-260
echo "== default s50: $(timeout 120 python scripts/profile_step.py 0 s50 512 quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"
