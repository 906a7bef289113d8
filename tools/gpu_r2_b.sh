#!/bin/bash
# round 2, GPU call B: GPU tests, cell-list rates with the compact kernel, launch list + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/b_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 300 python tools/bench_configs.py cfg3 cfg5 > gpurun_out/b_configs.jsonl 2> gpurun_out/b_configs.err
for t in "cipt=2" "cchunk=4" "cchunk=16" "cws=24" "cws=48" "cws=192" "cws=400"; do
  MDH_TUNE=$t timeout 200 python tools/bench_configs.py cfg3 > "gpurun_out/b_cfg3_$t.jsonl" 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches_cfg3.csv python tools/bench_configs.py cfg3 > gpurun_out/b_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 4 -c 1 -o gpurun_out/b_cellpair python tools/bench_configs.py cfg3 > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -20
