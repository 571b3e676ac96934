#!/bin/bash
# round-2 development call: new fused tests, variant tests, pipeline tests, micro-bench, pass sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
o=gpurun_out/r2_${TAG:-dev}
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_variants.py tests/test_gpu_pipeline.py tests/test_gpu_head_metrics.py -x -q -m gpu ) > ${o}_tests.log 2>&1
tail -n 15 ${o}_tests.log
timeout 300 python tools/kbench_fused.py > ${o}_kbench_fused.txt 2>&1
cat ${o}_kbench_fused.txt
timeout 300 python tools/pass_time.py > ${o}_pass_time.txt 2>&1
cat ${o}_pass_time.txt
