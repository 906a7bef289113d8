#!/bin/bash
# one ncu --set full capture of the S(q) lattice kernel (a 16-frame launch of cfg4)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sq-frames-per-step 16 ${SQARGS}"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice -s 3 -c 1 -f -o gpurun_out/${OUT:-r01_sq_mma} $B > gpurun_out/ncu_sq.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu_sq.log
