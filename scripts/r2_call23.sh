#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -m gpu -q -x -s -k "long_traces or config4" > gpurun_out/r2c23_tests.txt 2>&1; tail -8 gpurun_out/r2c23_tests.txt | cut -c1-200
