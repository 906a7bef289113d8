#!/bin/bash
# iteration session: GPU parity tests + bench variants (args: list of "tune:hist")
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests.log
for vh in "$@"; do
  v=${vh%%:*}; h=${vh##*:}
  MDH_TUNE="$v" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --hist $h > gpurun_out/bench_${v}_${h}.json 2> gpurun_out/bench_${v}_${h}.err
  echo "$v $h rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${v}_${h}.json"))
print(" evals/s %.3e  binned/s %.3e  e2e %.3e  frac %.3f  ms/step %.2f  clk %s" % (d["pairs_evaluated_per_s"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]))
PY
done
