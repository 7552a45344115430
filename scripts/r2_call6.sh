#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c6_tests.txt
echo "== tests =="; grep -v "^E   \s*+" gpurun_out/r2c6_tests.txt | tail -25 | cut -c1-200
for cfg in "A_default" "G_dxcluster PPB_REC_DX_CLUSTER=1"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 100 python scripts/profile_step.py 0 s50 512 > gpurun_out/r2c6_prof_$name.txt 2>&1
  echo "== $name: $(head -1 gpurun_out/r2c6_prof_$name.txt)"; sed -n 4,16p gpurun_out/r2c6_prof_$name.txt | cut -c1-130
done
timeout 100 python scripts/profile_step.py 0 gum 256 > gpurun_out/r2c6_prof_gum.txt 2>&1; head -1 gpurun_out/r2c6_prof_gum.txt
timeout 300 python bench.py > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
echo "== bench =="; cut -c1-300 gpurun_out/r2c6_bench.json; tail -4 gpurun_out/r2c6_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_score2|k_categorical|k_mixture|k_partials|k_finalize|k_normal' \
  -s 9 -c 12 -o gpurun_out/r2_scoring python scripts/ncu_scoring.py > gpurun_out/r2_ncu_scoring.log 2>&1
ncu -i gpurun_out/r2_scoring.ncu-rep --page raw --csv > gpurun_out/r2_scoring_raw.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_gum.csv \
  python scripts/ncu_step.py gum 3 > gpurun_out/r2_ncu_gum.log 2>&1
tail -2 gpurun_out/r2_ncu_scoring.log; tail -2 gpurun_out/r2_ncu_gum.log
