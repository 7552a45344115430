#!/usr/bin/env bash
# 8 GPUs: N = 1, 2, 4, 8 bench lines back to back (the driver's SCALE run) + dp tests at 8 ranks
set -u
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
  echo "== bench $n =="; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_bench_${n}gpu.json') if l.startswith('{')][-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('dp_step_phases_us_rank0'), d['ms_per_step_stats_rank0'], d['e2e'])
except Exception as e:
    print('no line', e)
PY
  tail -3 gpurun_out/r2_bench_${n}gpu.err
done
timeout 200 python bench.py --no-extra --cpu-budget 2 > gpurun_out/r2_bench_1gpu_scale.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r2_bench_1gpu_scale.json'));print(1,d['value'],d['ms_per_step'])"
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_dp8_tests.txt; cat gpurun_out/r2_dp8_tests.txt
