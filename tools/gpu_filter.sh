#!/bin/bash
# fp32-filter iteration: GPU parity tests (rdf), then the bench under several MDH_TUNE settings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rdf.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/tests_rdf.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_rdf.log
for v in "$@"; do
  MDH_TUNE="$v" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_${v}.json 2> gpurun_out/bench_${v}.err
  echo "$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${v}.json"))
print(" evals/s %.3e  binned/s %.3e  e2e %.3e  frac %.3f  ms/step %.2f launch_ms %.3f clk %s" % (d["pairs_evaluated_per_s"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["launch_ms"], d["clocks"]))
PY
done
