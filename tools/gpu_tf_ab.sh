#!/bin/bash
# fused tissue histogram in the row-form lift+argmax: parity tests, stand-alone timings, the pass with / without it
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_fused.py tests/test_gpu_exchange.py tests/test_gpu_pipeline.py -q -m gpu -x ) > gpurun_out/tf_tests.log 2>&1
tail -n 5 gpurun_out/tf_tests.log
timeout 300 python tools/kbench_fused.py 2>&1 | head -10 > gpurun_out/tf_kbench.txt
cat gpurun_out/tf_kbench.txt
( for a in 0 1 0 1; do echo "== TF=$a"; TF=$a NFLY=1,3 TMA=11 timeout 300 python tools/pass_overlap.py; done ) > gpurun_out/tf_pass_ab.txt 2>&1
cat gpurun_out/tf_pass_ab.txt
