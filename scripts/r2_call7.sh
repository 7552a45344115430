#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_grouped|k_cluster' -s 14 -c 7 \
  -o gpurun_out/r2_gemm_gum python scripts/ncu_step.py gum 4 > gpurun_out/r2_ncu_gemm_gum.log 2>&1
ncu -i gpurun_out/r2_gemm_gum.ncu-rep --page raw --csv > gpurun_out/r2_gemm_gum_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_cluster' -s 60 -c 3 \
  -o gpurun_out/r2_gemm_s50_fwd python scripts/ncu_step.py s50 2 > gpurun_out/r2_ncu_gemm_s50_fwd.log 2>&1
ncu -i gpurun_out/r2_gemm_s50_fwd.ncu-rep --page raw --csv > gpurun_out/r2_gemm_s50_fwd_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_grouped<1, 1>' -s 55 -c 4 \
  -o gpurun_out/r2_gemm_s50_bwd python scripts/ncu_step.py s50 2 > gpurun_out/r2_ncu_gemm_s50_bwd.log 2>&1
ncu -i gpurun_out/r2_gemm_s50_bwd.ncu-rep --page raw --csv > gpurun_out/r2_gemm_s50_bwd_raw.csv 2>/dev/null
tail -2 gpurun_out/r2_ncu_gemm_gum.log gpurun_out/r2_ncu_gemm_s50_fwd.log gpurun_out/r2_ncu_gemm_s50_bwd.log
rm -f gpurun_out/r2_scoring.ncu-rep
timeout 300 python bench.py > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c7_bench.json'));print(d['value'],d['e2e']['value']);print({k:v.get('value') for k,v in d['workloads'].items()})"
