#!/bin/bash
# round profile: launch list of the default bench command, then full-set captures of the
# hot kernels.  Each ncu run follows a plain run of the same command (exit 0).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
B1="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames-per-step 20 --sq-frames-per-step 16"
$B1 > gpurun_out/plain_b1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rdf_filter_kernel -s 3 -c 1 -f -o gpurun_out/r01_filter $B1 > gpurun_out/ncu_filter.log 2>&1
echo "filter rc=$?"
B2="$B1 --arith off"
$B2 > gpurun_out/plain_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rdf_allpairs -s 3 -c 1 -f -o gpurun_out/r01_pair $B2 > gpurun_out/ncu_pair.log 2>&1
echo "pair rc=$?"
$B1 > gpurun_out/plain_b1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice -s 3 -c 1 -f -o gpurun_out/r01_sq $B1 > gpurun_out/ncu_sq.log 2>&1
echo "sq rc=$?"
C="python tools/bench_configs.py cfg3"
$C > gpurun_out/plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rdf_cells_filter -s 2 -c 1 -f -o gpurun_out/r01_cells $C > gpurun_out/ncu_cells.log 2>&1
echo "cells rc=$?"
ls -la gpurun_out/*.ncu-rep
