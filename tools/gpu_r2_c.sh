#!/bin/bash
# round 2, GPU call C: cell-list kernel after pipeline changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c_smoke.log
timeout 900 python -m pytest tests/test_gpu_rdf.py -m gpu -q -x > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
for t in "" "cdbg=1" "cipt=2" "cipt=2,cchunk=4" "cchunk=4" "cchunk=2" "cws=24" "cws=192"; do
  MDH_TUNE=$t timeout 200 python tools/cells_speed.py >> gpurun_out/c_speed.jsonl 2>> gpurun_out/c_speed.err
done
timeout 300 python tools/bench_configs.py cfg3 cfg5 > gpurun_out/c_configs.jsonl 2> gpurun_out/c_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python tools/cells_speed.py > gpurun_out/c_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 3 -c 1 -o gpurun_out/c_cellpair python tools/cells_speed.py cfg3 > gpurun_out/c_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 3 -c 1 -o gpurun_out/c_cellpair5 python tools/cells_speed.py cfg5 > gpurun_out/c_ncu3.log 2>&1
ls -la gpurun_out | tail -12
