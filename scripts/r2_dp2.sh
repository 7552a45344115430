#!/usr/bin/env bash
# 2 GPUs: data-parallel parity tests (kept log for profiles/) + the 2-GPU bench line
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_dp2_gpus.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -v 2>&1 | tail -20 > gpurun_out/r2_dp2_tests.txt
echo "== dp tests =="; cat gpurun_out/r2_dp2_tests.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "== bench 2 =="; cut -c1-700 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
