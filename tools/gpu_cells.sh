#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rdf.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/tests_rdf.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_rdf.log
timeout 900 python tools/bench_configs.py cfg1 cfg3 cfg5 2> gpurun_out/configs.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], '| e2e frames/s %.1f' % d['e2e_frames_per_s'], '| kernel_ms %.3f' % d['kernel_ms'], '| evals/s %.3e' % (d.get('kernel_evaluations_per_s') or 0), '| binned/s e2e %.3e' % (d.get('e2e_pairs_binned_per_s') or 0), '| frac %.3f' % d['fp64_pipe_frac'])
"
tail -3 gpurun_out/configs.err
