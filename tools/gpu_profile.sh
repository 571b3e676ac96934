#!/bin/bash
# profiling evidence in one gpurun call: ncu --set full over the hot kernels (tools/prof_r2.py, last = warm launch of
# each kernel is summarised), compute-sanitizer memcheck + racecheck over the fused-kernel tests and the smoke pass
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kernel' -f -o gpurun_out/r02_full python tools/prof_r2.py > gpurun_out/r02_ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_fused.py -q -m gpu -x > gpurun_out/r02_sanitizer_memcheck_fused.log 2>&1
echo "memcheck fused rc=$?" >> gpurun_out/r02_sanitizer_memcheck_fused.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/prof_r2.py > gpurun_out/r02_sanitizer_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?" >> gpurun_out/r02_sanitizer_racecheck_kernels.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?" >> gpurun_out/r02_sanitizer_memcheck_smoke.log
tail -n 3 gpurun_out/r02_ncu_full.log gpurun_out/r02_sanitizer_*.log
ls -la gpurun_out/r02_full.ncu-rep
