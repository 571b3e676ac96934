#!/bin/bash
# second GPU call of the session: the 2-D map kernels and the DIFF lift+argmax variants
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -m pytest tests/test_gpu_multimodal.py tests/test_gpu_sampler.py -q -m gpu ) > gpurun_out/c2_new_tests.log 2>&1
for v in 4 5; do
  ( LDIFF_ARGMAX_VARIANT=$v python -m pytest tests/test_gpu_head_metrics.py tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -q -m gpu ) > gpurun_out/c2_tests_variant$v.log 2>&1
done
for v in 0 4 5 1; do
  LDIFF_ARGMAX_VARIANT=$v python tools/kbench_argmax.py >> gpurun_out/c2_argmax.txt 2>&1
  TAG="argmax variant $v" LDIFF_ARGMAX_VARIANT=$v python tools/pass_sched.py >> gpurun_out/c2_pass.txt 2>&1
done
python tools/kbench.py > gpurun_out/c2_kbench.txt 2>&1
tail -2 gpurun_out/c2_new_tests.log gpurun_out/c2_tests_variant4.log gpurun_out/c2_tests_variant5.log
cat gpurun_out/c2_argmax.txt gpurun_out/c2_pass.txt
grep -E "laplace_map|scaled_residual|laplace_big" gpurun_out/c2_kbench.txt
