#!/usr/bin/env bash
# First GPU call of a round that inherits opt-in code: run the whole GPU suite including the tests of code that
# has never run on hardware, then A/B the opt-in kernels against the defaults.  One gpurun call, ~3 GPU-minutes:
#
#   gpurun --timeout 600 -- 'bash scripts/first_hardware_run.sh'
#
# Outputs land in gpurun_out/ (merged back by gpurun): tests_default.txt, tests_unvalidated.txt, ab_optin.json
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests_default.txt
echo "== default suite =="; tail -3 gpurun_out/tests_default.txt
PPB_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests -m gpu -q -k "offline or optim or staged or fused_cell or infer_golden" 2>&1 \
  | tail -40 > gpurun_out/tests_unvalidated.txt
echo "== opt-in suite =="; tail -12 gpurun_out/tests_unvalidated.txt
timeout 200 python scripts/ab_optin.py > gpurun_out/ab_optin.json 2> gpurun_out/ab_optin.err
echo "== A/B =="; cat gpurun_out/ab_optin.json; tail -3 gpurun_out/ab_optin.err
