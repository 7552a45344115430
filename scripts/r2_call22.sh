#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c22_tests.txt 2>&1; tail -3 gpurun_out/r2c22_tests.txt | cut -c1-200
run() { name=$1; shift; for cfg in "gum 256" "s50 512"; do echo "== $name $cfg: $(env "$@" timeout 120 python scripts/profile_step.py 0 $cfg quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done; }
run A_default PPB_X=0
run B_allsplits PPB_SKIP_EMPTY_SPLITS=0
run A2_default PPB_X=0
run B2_allsplits PPB_SKIP_EMPTY_SPLITS=0
