#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_grouped_persistent' -s 4 -c 2 -o gpurun_out/r2j_persist python scripts/gemm_sat.py > gpurun_out/r2j_ncu_persist.log 2>&1
ncu -i gpurun_out/r2j_persist.ncu-rep --page raw --csv > gpurun_out/r2j_persist_raw.csv 2>/dev/null
rm -f gpurun_out/r2j_persist.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/r2j_launches_s50.csv python scripts/ncu_step.py s50 2 > gpurun_out/r2j_ncu_s50.log 2>&1
tail -2 gpurun_out/r2j_ncu_persist.log; ls -la gpurun_out/r2j_* | cut -c20-100
