#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2_baseline_shapes.txt
echo "== baseline shapes =="; tail -25 gpurun_out/r2_baseline_shapes.txt
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_tests_all.txt
echo "== all =="; tail -5 gpurun_out/r2_tests_all.txt
timeout 120 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2_prof_gum.txt 2>&1
timeout 120 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2_prof_s50.txt 2>&1
head -45 gpurun_out/r2_prof_gum.txt | cut -c1-150
