#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
run() { name=$1; shift; for cfg in "gum 256"; do echo "== $name $cfg: $(env "$@" timeout 120 python scripts/profile_step.py 0 $cfg quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done; }
run A_pdl1 PPB_PDL=1
run B_pdl2 PPB_PDL=2
run C_pdl3 PPB_PDL=3
run D_pdl0 PPB_PDL=0
run A2_pdl1 PPB_PDL=1
run B2_pdl2 PPB_PDL=2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c19_tests.txt 2>&1; tail -3 gpurun_out/r2c19_tests.txt
