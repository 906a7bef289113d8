#!/bin/bash
mkdir -p gpurun_out
C="python tools/bench_configs.py cfg3"
$C > gpurun_out/plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rdf_cells_filter -s 2 -c 1 -f -o gpurun_out/r01_cells $C > gpurun_out/ncu_cells.log 2>&1
echo "cells rc=$?"
