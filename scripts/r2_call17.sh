#!/usr/bin/env bash
# final-evidence call: tests, bench (both arms), CUPTI kernel times without PDL, ncu launch lists and --set full captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c17_tests.txt 2>&1; tail -3 gpurun_out/r2c17_tests.txt
timeout 400 python bench.py > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c17_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()});print([(x['kernel'],round(x['frac_of_hbm'],3)) for x in d['extra']['scoring_hbm_roofline']]);print(d['extra']['gate_gemm_saturating_4096x2048x512'])"
timeout 300 python bench.py --impl reference > gpurun_out/r2c17_bench_ref.json 2> gpurun_out/r2c17_bench_ref.err; cut -c1-160 gpurun_out/r2c17_bench_ref.json
PPB_PDL=0 timeout 200 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c17_cupti_s50_nopdl.txt 2>&1; head -14 gpurun_out/r2c17_cupti_s50_nopdl.txt | cut -c1-150
PPB_PDL=0 timeout 200 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c17_cupti_gum_nopdl.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_gum.csv python scripts/ncu_step.py gum 3 > gpurun_out/r2f_ncu_gum.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/r2f_launches_s50.csv python scripts/ncu_step.py s50 2 > gpurun_out/r2f_ncu_s50.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_cluster' -s 60 -c 3 -o gpurun_out/r2f_lstm_fwd python scripts/ncu_step.py s50 2 > gpurun_out/r2f_ncu_lstm.log 2>&1
ncu -i gpurun_out/r2f_lstm_fwd.ncu-rep --page raw --csv > gpurun_out/r2f_lstm_fwd_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_grouped|k_cluster' -s 10 -c 8 -o gpurun_out/r2f_gemm_gum python scripts/ncu_step.py gum 4 > gpurun_out/r2f_ncu_gemm_gum.log 2>&1
ncu -i gpurun_out/r2f_gemm_gum.ncu-rep --page raw --csv > gpurun_out/r2f_gemm_gum_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_score2|k_categorical|k_mixture|k_partials|k_finalize|k_normal' -s 9 -c 12 -o gpurun_out/r2f_scoring python scripts/ncu_scoring.py > gpurun_out/r2f_ncu_scoring.log 2>&1
ncu -i gpurun_out/r2f_scoring.ncu-rep --page raw --csv > gpurun_out/r2f_scoring_raw.csv 2>/dev/null
rm -f gpurun_out/r2f_scoring.ncu-rep gpurun_out/r2f_gemm_gum.ncu-rep
ls -la gpurun_out/r2f_* | cut -c20-120
