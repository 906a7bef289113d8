#!/bin/bash
# one ncu --set full capture of the S(q) lattice kernel (cfg4, ${FPS:-128} frames per launch)
mkdir -p gpurun_out
B="python tools/sq_speed.py ${SQK:-lattice_dmma} ${FPS:-128}"
$B > gpurun_out/plain_sq.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice -s 8 -c 1 -f -o gpurun_out/${OUT:-r01_sq_mma} $B > gpurun_out/ncu_sq.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_sq.log
