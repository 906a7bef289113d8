#!/bin/bash
# round 2, GPU call E2: (repeat of E with the S(q) idle-warp fix) all GPU tests with a per-test
# timeout; cell-pair block shapes; S(q) after the spill fix; overhead + H2D probes; bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python tools/sq_speed.py lattice_dmma 128 16 > gpurun_out/e_sq.log 2>&1 || echo "sq_speed rc=$?" >> gpurun_out/e_sq.log
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout 100 python tools/h2d_probe.py > gpurun_out/e_h2d.json 2>&1
for lib in "" t224b2 t160b3; do
  for t in "" "cipt=2"; do
    if [ -n "$lib" ]; then export MDH_B200_LIB=$PWD/mdhelper_b200/libmdh_b200_$lib.so; else unset MDH_B200_LIB; fi
    echo "{\"lib\": \"$lib\"}" >> gpurun_out/e_speed.jsonl
    MDH_TUNE=$t timeout 100 python tools/cells_speed.py >> gpurun_out/e_speed.jsonl 2>> gpurun_out/e_speed.err
  done
done
unset MDH_B200_LIB
MDH_TUNE="cdbg=1" timeout 100 python tools/cells_speed.py cfg3 > gpurun_out/e_dbg.log 2>&1
for nm in 20 32 10; do timeout 100 python tools/sq_speed.py lattice_dmma 128 $nm >> gpurun_out/e_sq.log 2>&1; done
timeout 200 python tools/overhead_probe.py > gpurun_out/e_overhead.json 2> gpurun_out/e_overhead.err
timeout 500 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sq_lattice_mma -s 3 -c 1 -o gpurun_out/e_sq python tools/sq_speed.py lattice_dmma 128 16 > gpurun_out/e_ncu_sq.log 2>&1
ls -la gpurun_out | tail -8
