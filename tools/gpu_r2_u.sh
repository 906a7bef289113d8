#!/bin/bash
# round 2, GPU call U: filter kernel, 4-stage vs 3-stage pipeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, lib, tune
  if [ -n "$2" ]; then export MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_$2.so; else unset MDH_B200_LIB; fi
  MDH_TUNE="$3" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-secondary > gpurun_out/u_$1.json 2> gpurun_out/u_$1.err
  python -c "
import json
d=json.loads(open('gpurun_out/u_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'])"
}
run p4 p4 ""
run p3 "" ""
run p4_b p4 ""
run p3_b "" ""
MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_p4.so timeout 300 python -m pytest tests/test_gpu_rdf.py -m gpu -q --timeout 150 -x -k "not cells and not triclinic" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
tail -3 gpurun_out/u_pytest.log
