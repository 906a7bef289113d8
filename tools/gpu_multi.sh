#!/bin/bash
# multi-GPU: bench under torchrun on N GPUs (arg 1), plus microbench and single-GPU bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 200 > gpurun_out/mb_clocks.csv &
SMI_PID=$!; sleep 0.3
./tools/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err; echo "microbench rc=$?"
kill $SMI_PID 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; tail -2 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
for n in ("n1", "n$N", "ref"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        s=d.get("secondary") or {}
        print(n, "value %.3e e2e %.3e ms/step %.2f frac %s | sq %s e2e %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], (d.get("roofline") or {}).get("frac"), s.get("value"), (s.get("e2e") or {}).get("value")))
    except Exception as e: print(n, "ERR", e)
PY
cat gpurun_out/microbench.json
