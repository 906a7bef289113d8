#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/bench_configs.py "$@" > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "rc=$?"; tail -5 gpurun_out/configs.err
cat gpurun_out/configs.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
s=d["secondary"]
print("value %.3e e2e %.3e ms/step %.2f launch_ms %.2f frac %.3f | sq %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["frac"], s["value"], s["e2e"]["value"], s["roofline"]["frac"]))
PY
