#!/bin/bash
# ncu: full-set capture of one pair kernel on a short bench run; args: kernel-regex tune
mkdir -p gpurun_out
K=${1:-rdf_filter}; T=${2:-ipt=4}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --frames-per-step 20"
MDH_TUNE=$T $B > gpurun_out/plain.log 2>&1 &&
MDH_TUNE=$T ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${K}_${T} $B > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu.log
