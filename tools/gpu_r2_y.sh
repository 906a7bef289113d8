#!/bin/bash
# round 2, GPU call Y: per-rank breakdown of the cfg4 strong-scaling pass on 8 GPUs, and of the same per-rank work alone
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROBE_FRAMES=125 timeout 120 python tools/strong_probe.py > gpurun_out/y_probe_n1_125.jsonl 2> gpurun_out/y_probe_n1.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/strong_probe.py > gpurun_out/y_probe_n8.jsonl 2> gpurun_out/y_probe_n8.err
cat gpurun_out/y_probe_n1_125.jsonl gpurun_out/y_probe_n8.jsonl | cut -c1-700
tail -3 gpurun_out/y_probe_n8.err | cut -c1-200
