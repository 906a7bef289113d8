#!/bin/bash
# round 2, GPU call S: cell-pair kernel with LOP3 addressing (tests, A/B against the previous build); full capture of the new filter kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rdf.py -m gpu -q --timeout 150 -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 200 python tools/cells_speed.py > gpurun_out/s_speed_new.jsonl 2>&1
MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_base.so timeout 200 python tools/cells_speed.py > gpurun_out/s_speed_base.jsonl 2>&1
timeout 200 python tools/cells_speed.py > gpurun_out/s_speed_new2.jsonl 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rdf_filter_kernel -s 4 -c 1 -o gpurun_out/s_filter python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/s_ncu.log 2>&1
tail -3 gpurun_out/s_pytest.log
cat gpurun_out/s_speed_*.jsonl | cut -c1-170
