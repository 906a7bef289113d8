#!/bin/bash
# round 2, last GPU call: all GPU tests, smoke and the default bench line on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/zz_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/zz_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/zz_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/zz_smoke.log
timeout 600 python bench.py > gpurun_out/zz_bench.json 2> gpurun_out/zz_bench.err; echo "bench rc=$?" >> gpurun_out/zz_bench.err
tail -3 gpurun_out/zz_pytest.log; tail -2 gpurun_out/zz_smoke.log; tail -2 gpurun_out/zz_bench.err; cut -c1-300 gpurun_out/zz_bench.json
