#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py::test_marsaglia_inference_compilation tests/test_network_gpu.py tests/test_optim_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2c5_tests.txt
echo "== tests =="; grep -v "^E   \s*+" gpurun_out/r2c5_tests.txt | tail -25 | cut -c1-200
PPB_FLAT_ADAM=1 timeout 200 python -m pytest tests/test_model_gpu.py::test_marsaglia_inference_compilation -m gpu -q 2>&1 | tail -8 | cut -c1-200
for cfg in "A_default" "B_single PPB_SINGLE_STREAM=1" "C_nofusedbwd PPB_FUSED_CELL_BWD=0" "D_nocluster PPB_NO_CLUSTER=1" "E_single_nofusedbwd PPB_SINGLE_STREAM=1 PPB_FUSED_CELL_BWD=0" "F_single_legacy PPB_SINGLE_STREAM=1 PPB_NO_CLUSTER=1 PPB_FUSED_CELL=0"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 100 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c5_prof_$name.txt 2>&1
  echo "== $name: $(head -1 gpurun_out/r2c5_prof_$name.txt)"; sed -n 4,11p gpurun_out/r2c5_prof_$name.txt | cut -c1-130
done
PPB_HOST_STEP_GRAPH=1 timeout 200 python bench.py --no-extra --cpu-budget 1 > gpurun_out/r2c5_bench_hostgraph.json 2> gpurun_out/r2c5_bench_hostgraph.err
echo "== host graph =="; python -c "
import json;d=json.load(open('gpurun_out/r2c5_bench_hostgraph.json'));print(d['ms_per_step'], d['e2e'])"; tail -2 gpurun_out/r2c5_bench_hostgraph.err
PPB_HOST_STEP_GRAPH=1 timeout 200 python -m pytest tests/test_host_step_gpu.py -m gpu -q 2>&1 | tail -5
