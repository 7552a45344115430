#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c3_tests.txt
echo "== tests =="; tail -6 gpurun_out/r2c3_tests.txt
timeout 400 python bench.py > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "== bench =="; cut -c1-1500 gpurun_out/r2c3_bench.json; tail -3 gpurun_out/r2c3_bench.err
PPB_SINGLE_STREAM=1 timeout 200 python bench.py --no-extra --cpu-budget 1 > gpurun_out/r2c3_bench_single.json 2>/dev/null
echo "== single stream =="; cut -c1-400 gpurun_out/r2c3_bench_single.json
timeout 120 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c3_prof_gum.txt 2>&1
head -40 gpurun_out/r2c3_prof_gum.txt | cut -c1-150
