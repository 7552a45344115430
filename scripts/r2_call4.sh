#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c4_tests_gemm.txt
echo "== gemm tests =="; tail -6 gpurun_out/r2c4_tests_gemm.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c4_tests.txt
echo "== tests =="; tail -12 gpurun_out/r2c4_tests.txt
timeout 120 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c4_prof_gum.txt 2>&1
head -34 gpurun_out/r2c4_prof_gum.txt | cut -c1-150
timeout 120 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c4_prof_s50.txt 2>&1
head -30 gpurun_out/r2c4_prof_s50.txt | cut -c1-150
timeout 420 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "== bench =="; cut -c1-600 gpurun_out/r2c4_bench.json; tail -25 gpurun_out/r2c4_bench.err
