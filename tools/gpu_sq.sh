#!/bin/bash
# S(q) session: GPU parity tests + bench with the DMMA and the scalar-DFMA lattice kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sq.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/tests_sq.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/tests_sq.log
for k in ${SQK:-lattice_dmma lattice_fp64}; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sq-kernel $k > gpurun_out/bench_sq_$k.json 2> gpurun_out/bench_sq_$k.err
  echo "bench rc=$? kernel=$k"; tail -3 gpurun_out/bench_sq_$k.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sq_$k.json"))
s=d["secondary"]
print("RDF evals/s %.3e frac %.3f e2e %.3e | SQ frames/s %.1f e2e %.1f frac %.3f ms/step %.3f launch_ms %.3f" % (d["pairs_evaluated_per_s"], d["roofline"]["frac"], d["e2e"]["value"], s["value"], s["e2e"]["value"], s["roofline"]["frac"], s["ms_per_step"], s["roofline"]["launch_ms"]))
PY
done
