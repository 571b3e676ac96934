#!/bin/bash
# row-form lift+argmax: parity tests, stand-alone timings, and the pass with the row / column form (same box)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_fused.py tests/test_gpu_head_metrics.py tests/test_gpu_pipeline.py tests/test_gpu_edge_cases.py -q -m gpu -x ) > gpurun_out/row_tests.log 2>&1
tail -n 5 gpurun_out/row_tests.log
timeout 300 python tools/kbench_fused.py 2>&1 | head -12 > gpurun_out/row_kbench.txt
cat gpurun_out/row_kbench.txt
( for a in 1 0 1 0; do echo "== AMAX=$a"; AMAX=$a NFLY=1,3 TMA=11 timeout 300 python tools/pass_overlap.py; done ) > gpurun_out/row_pass_ab.txt 2>&1
cat gpurun_out/row_pass_ab.txt
