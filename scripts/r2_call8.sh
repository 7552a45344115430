#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for cfg in "A_default" "G4_dxcluster4 PPB_REC_DX_CLUSTER=4" "G2_dxcluster2 PPB_REC_DX_CLUSTER=2"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 100 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c8_prof_$name.txt 2>&1
  echo "== $name: $(head -1 gpurun_out/r2c8_prof_$name.txt)"; sed -n 4,10p gpurun_out/r2c8_prof_$name.txt | cut -c1-130
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/r2_launches_s50.csv \
  python scripts/ncu_step.py s50 2 > gpurun_out/r2_ncu_s50.log 2>&1
tail -1 gpurun_out/r2_ncu_s50.log
