#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/marsaglia_ess.py 2>&1 | grep -v "^Total addr\|^Creating" > gpurun_out/r2c10_marsaglia_ess.txt; cat gpurun_out/r2c10_marsaglia_ess.txt
timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c10_prof_s50.txt 2>&1; grep "tc launch" gpurun_out/r2c10_prof_s50.txt | sed -n '1,12p;60,70p' | cut -c1-250
timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c10_prof_gum.txt 2>&1; grep "tc launch" gpurun_out/r2c10_prof_gum.txt | cut -c1-250
timeout 300 python scripts/ab_optin.py mix > gpurun_out/r2c10_ab_mix.json 2>&1; cat gpurun_out/r2c10_ab_mix.json
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_model_gpu.py::test_marsaglia_inference_compilation > gpurun_out/r2c10_tests.txt 2>&1; tail -3 gpurun_out/r2c10_tests.txt
