#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice -s 3 -c 1 -f -o gpurun_out/sq_lattice $B > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu.log
