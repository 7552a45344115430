#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c20_tests.txt 2>&1; tail -8 gpurun_out/r2c20_tests.txt | cut -c1-200
run() { name=$1; shift; for cfg in "gum 256" "s50 512"; do echo "== $name $cfg: $(env "$@" timeout 120 python scripts/profile_step.py 0 $cfg quick 2>&1 | grep -v Warn | tr '\n' ' ' | cut -c1-200)"; done; }
run A_default PPB_X=0
run B_nofusehead PPB_FUSE_HEAD_OUT=0
run A2_default PPB_X=0
timeout 400 python bench.py > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2c20_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']);print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['workloads'].items()})"
