#!/bin/bash
# round 2, GPU call W: final state -- all GPU tests, smoke, the default bench line, launch list and full capture of the filter kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/w_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/w_smoke.log
timeout 900 python bench.py > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc=$?" >> gpurun_out/w_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/w_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/w_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_filter_kernel -s 4 -c 1 -o gpurun_out/w_filter python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/w_ncu.log 2>&1
tail -3 gpurun_out/w_pytest.log; tail -2 gpurun_out/w_smoke.log; tail -2 gpurun_out/w_bench.err; cut -c1-400 gpurun_out/w_bench.json
