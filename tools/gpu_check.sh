#!/bin/bash
# One-call GPU validation (run under gpurun): new parity tests first, then the whole GPU suite,
# smoke, the bench line, the per-kernel bench, a sanitizer pass over the new kernels.
# Every step writes under gpurun_out/ as it goes, so a call cut short keeps what it finished.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
( time python -m pytest tests/test_gpu_multimodal.py tests/test_dataset_tables.py tests/test_gpu_sampler.py -q -m gpu ) > gpurun_out/new_tests.log 2>&1
( time python -m pytest tests -x -q -m gpu --durations=12 ) > gpurun_out/gpu_tests.log 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python tools/kbench.py > gpurun_out/kbench.txt 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_multimodal.py -q -m gpu -k "noising or unnoise or bf16" > gpurun_out/sanitizer_memcheck_multimodal.log 2>&1
tail -3 gpurun_out/new_tests.log gpurun_out/gpu_tests.log gpurun_out/smoke.log gpurun_out/sanitizer_memcheck_multimodal.log
cat gpurun_out/bench_1gpu.json | cut -c1-400
