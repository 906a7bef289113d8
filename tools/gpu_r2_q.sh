#!/bin/bash
# round 2, GPU call Q: final test run + profiles (launch lists with DRAM bytes, full-set captures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
# launch list of the default bench (every kernel of a short run)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/q_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/q_ncu_bench.log 2>&1
# cell-list pipeline: duration + DRAM bytes of every kernel (cfg3, device-resident)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/q_launches_cells.csv python tools/cells_speed.py cfg3 > gpurun_out/q_ncu_cells.log 2>&1
MDH_TUNE=cws=8 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/q_launches_cells_ws8.csv python tools/cells_speed.py cfg3 > gpurun_out/q_ncu_cells8.log 2>&1
MDH_TUNE=cws=8 timeout 100 python tools/cells_speed.py cfg3 > gpurun_out/q_speed_ws8.jsonl 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rdf_cellpair -s 2 -c 1 -o gpurun_out/q_cellpair python tools/cells_speed.py cfg3 > gpurun_out/q_ncu2.log 2>&1
tail -3 gpurun_out/q_pytest.log
