#!/bin/bash
# round 2, GPU call D: per-cell overhead trims; block-shape variants of the cell-pair kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rdf.py -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
for lib in "" t192b2 t128b4 t128b3; do
  for t in "" "cipt=2" "cipt=2,cchunk=4" "cchunk=4,cws=192"; do
    if [ -n "$lib" ]; then export MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_$lib.so; else unset MDH_B200_LIB; fi
    echo "{\"lib\": \"$lib\"}" >> gpurun_out/d_speed.jsonl
    MDH_TUNE=$t timeout 200 python tools/cells_speed.py >> gpurun_out/d_speed.jsonl 2>> gpurun_out/d_speed.err
  done
done
unset MDH_B200_LIB
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 2 -c 1 -o gpurun_out/d_cellpair python tools/cells_speed.py cfg3 > gpurun_out/d_ncu2.log 2>&1
MDH_TUNE="cipt=2" timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 2 -c 1 -o gpurun_out/d_cellpair_ipt2 python tools/cells_speed.py cfg3 > gpurun_out/d_ncu3.log 2>&1
ls -la gpurun_out | tail -8
