#!/bin/bash
# round-end session on one GPU: all GPU parity tests, the default bench line, its ncu launch
# list, and one full-set capture of the S(q) DMMA kernel (a 128-frame launch of cfg4)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain_b1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice -s 3 -c 1 -f -o gpurun_out/r01_sq $B > gpurun_out/ncu_sq.log 2>&1
echo "sq rc=$?"
