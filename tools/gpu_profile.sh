#!/bin/bash
# profiling evidence in one gpurun call: compute-sanitizer racecheck / memcheck over the hot kernels and the fused-kernel
# tests, then ncu --set full over the warm iteration of tools/prof_r2.py; the report is summarised and exported ON THE
# BOX (raw page + the source pages of the two heaviest kernels) because the .ncu-rep itself is over gpurun's 64 MiB
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/prof_r2.py > gpurun_out/r02_sanitizer_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?" >> gpurun_out/r02_sanitizer_racecheck_kernels.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_fused.py -q -m gpu -x > gpurun_out/r02_sanitizer_memcheck_fused.log 2>&1
echo "memcheck fused rc=$?" >> gpurun_out/r02_sanitizer_memcheck_fused.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?" >> gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'kernel' -f -o /tmp/r02_full python tools/prof_r2.py > gpurun_out/r02_ncu_full.log 2>&1
python tools/summarize_ncu.py /tmp/r02_full.ncu-rep gpurun_out/r02_ncu_full_kernels > gpurun_out/r02_ncu_summarize.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw.csv 2>/dev/null
ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:'lift_argmax_env' > gpurun_out/r02_ncu_source_lift_argmax_env.csv 2>/dev/null
ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:'decode_tail_tma' > gpurun_out/r02_ncu_source_decode_tail_tma.csv 2>/dev/null
ncu -i /tmp/r02_full.ncu-rep --page details -k regex:'lift_argmax_env|decode_tail_tma|lut_paint_hist|head_logits_tma' > gpurun_out/r02_ncu_details.txt 2>/dev/null
gzip -f gpurun_out/r02_ncu_source_*.csv gpurun_out/r02_ncu_raw.csv
ls -la /tmp/r02_full.ncu-rep gpurun_out/ | tail -20
tail -n 3 gpurun_out/r02_ncu_full.log gpurun_out/r02_sanitizer_*.log
grep -B2 -A12 "hazard" gpurun_out/r02_sanitizer_racecheck_kernels.log | head -80
