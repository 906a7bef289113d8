#!/bin/bash
# round 2, GPU call R: float64-coordinate S(q)/ISF paths; column histogram of the filter kernel (A/B against the previous build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/r_bench_new.json 2> gpurun_out/r_bench_new.err
MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_base.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/r_bench_base.json 2> gpurun_out/r_bench_base.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/r_bench_new2.json 2>> gpurun_out/r_bench_new.err
tail -3 gpurun_out/r_pytest.log
for f in gpurun_out/r_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'])"; done
