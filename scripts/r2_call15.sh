#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvcc -arch=sm_100a -o /tmp/cluster_occ scripts/cluster_occupancy.cu && /tmp/cluster_occ | tee gpurun_out/r2c15_cluster_occupancy.txt
run() { name=$1; shift; for cfg in "gum 256" "s50 512"; do set -- "$@"; echo "== $name $cfg: $(env "$@" timeout 120 python scripts/profile_step.py 0 $cfg quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done; }
run A_default PPB_X=0
run B_pdl1 PPB_PDL=1
run C_pdl0 PPB_PDL=0
run D_noprefetch PPB_LSTM_B_PREFETCH=0
run E_dxcluster4 PPB_REC_DX_CLUSTER=4
run F_dxcluster8 PPB_REC_DX_CLUSTER=8
run A2_default PPB_X=0
PPB_REC_DX_CLUSTER=4 timeout 600 python -m pytest tests -m gpu -q -x -k "network or baseline or fused or pdl or host_step" > gpurun_out/r2c15_tests_dx4.txt 2>&1; tail -3 gpurun_out/r2c15_tests_dx4.txt
