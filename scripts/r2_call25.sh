#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x -k "test_gemm_packed and not tn and not cluster" > gpurun_out/r2c25_gemm.txt 2>&1; tail -4 gpurun_out/r2c25_gemm.txt | cut -c1-200
PPB_PERSISTENT=0 timeout 100 python scripts/gemm_sat.py 2>&1 | tail -1
PPB_PERSISTENT=1 timeout 100 python scripts/gemm_sat.py 2>&1 | tail -1
PPB_PERSISTENT=1 timeout 600 python -m pytest tests -m gpu -q -x -k "baseline or network or fused_cell or pdl" > gpurun_out/r2c25_net.txt 2>&1; tail -4 gpurun_out/r2c25_net.txt | cut -c1-200
for v in 0 1; do echo "== persistent=$v s50: $(PPB_PERSISTENT=$v timeout 120 python scripts/profile_step.py 0 s50 512 quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done
