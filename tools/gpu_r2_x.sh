#!/bin/bash
# round 2, final multi-GPU lines: the default bench (weak headline + strong passes with the NCCL parity check) on N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
MDH_BENCH_STALL_S=120 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/x_bench_n$N.json 2> gpurun_out/x_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/x_bench_n$N.err
grep -v "Warning\|warn" gpurun_out/x_bench_n$N.err | tail -8 | cut -c1-250
cut -c1-300 gpurun_out/x_bench_n$N.json
