#!/bin/bash
# first GPU session: microbenchmarks, smoke, GPU parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 ./tools/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err
echo "microbench rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -40 gpurun_out/tests.log
cat gpurun_out/microbench.json
