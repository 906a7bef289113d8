#!/bin/bash
# round 2, GPU call A: smoke, GPU tests, cell-list rates, default bench, launch list + ncu of the cell-pair kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python tools/bench_configs.py cfg3 cfg5 > gpurun_out/a_configs.jsonl 2> gpurun_out/a_configs.err
for t in "cipt=2" "cchunk=4" "cchunk=16" "cws=12" "cws=96"; do
  MDH_TUNE=$t timeout 200 python tools/bench_configs.py cfg3 > "gpurun_out/a_cfg3_$t.jsonl" 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches_cfg3.csv python tools/bench_configs.py cfg3 > gpurun_out/a_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 4 -c 1 -o gpurun_out/a_cellpair python tools/bench_configs.py cfg3 > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out | tail -20
