#!/bin/bash
# round 2, GPU call V: cell-pair row loop as a software pipeline (tests, A/B against the previous loop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/cells_speed.py > gpurun_out/v_speed_new.jsonl 2>&1
MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_c2.so timeout 200 python tools/cells_speed.py > gpurun_out/v_speed_c2.jsonl 2>&1
MDH_TUNE=cipt=2 timeout 200 python tools/cells_speed.py > gpurun_out/v_speed_new_ipt2.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_rdf.py -m gpu -q --timeout 150 -x > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -3 gpurun_out/v_pytest.log
cat gpurun_out/v_speed_*.jsonl | cut -c1-150
