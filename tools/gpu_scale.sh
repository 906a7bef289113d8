#!/bin/bash
# scaling check: bench under torchrun on N GPUs (arg 1), no CPU baseline
N=${1:-8}
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
s=d.get("secondary") or {}
print("n$N value %.3e e2e %.3e ms/step %.2f frac %s | sq %s e2e %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], (d.get("roofline") or {}).get("frac"), s.get("value"), (s.get("e2e") or {}).get("value")))
PY
