#!/bin/bash
# final state of the round: whole GPU suite, smoke, bench line; then (time permitting) racecheck of the persistent kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/c5_gpu_tests.log 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/c5_smoke.log 2>&1
python bench.py > gpurun_out/c5_bench_1gpu.json 2> gpurun_out/c5_bench_1gpu.err
LDIFF_ARGMAX_PERSIST=52 timeout 70 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_head_metrics.py -q -m gpu -k "lift_argmax_bit_exact or ties" > gpurun_out/c5_racecheck_persist.log 2>&1
for f in gpurun_out/c5_gpu_tests.log gpurun_out/c5_smoke.log gpurun_out/c5_racecheck_persist.log; do tail -n 4 $f; done
cut -c1-220 gpurun_out/c5_bench_1gpu.json
