#!/bin/bash
# round 2, GPU call N: after the stream fix -- all GPU tests, full bench, cell-list and S(q) rates
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
MDH_BENCH_STALL_S=120 timeout 400 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?" >> gpurun_out/n_bench.err
timeout 100 python tools/cells_speed.py > gpurun_out/n_speed.jsonl 2> gpurun_out/n_speed.err
MDH_TUNE=cipt=2 timeout 100 python tools/cells_speed.py >> gpurun_out/n_speed.jsonl 2>> gpurun_out/n_speed.err
timeout 100 python tools/bench_configs.py cfg3 cfg5 > gpurun_out/n_configs.jsonl 2> gpurun_out/n_configs.err
timeout 100 python tools/overhead_probe.py > gpurun_out/n_overhead.json 2> gpurun_out/n_overhead.err
tail -3 gpurun_out/n_pytest.log; tail -12 gpurun_out/n_bench.err
