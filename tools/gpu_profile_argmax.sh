#!/bin/bash
# evidence for the row-form lift+argmax and the uint16 paint: sanitizer runs and a warm ncu --set full capture
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_variants.py tests/test_gpu_fused.py -q -m gpu -x -k "lift_argmax or envelope or uint16 or tall_bands" > gpurun_out/r02b_sanitizer_memcheck_argmax.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02b_sanitizer_memcheck_argmax.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/prof_argmax.py > gpurun_out/r02b_sanitizer_racecheck_argmax.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02b_sanitizer_racecheck_argmax.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'lift_argmax|lut_paint' -f -o gpurun_out/r02b_argmax python tools/prof_argmax.py > gpurun_out/r02b_ncu_argmax.log 2>&1
python tools/summarize_ncu.py gpurun_out/r02b_argmax.ncu-rep gpurun_out/r02b_ncu_argmax_kernels > gpurun_out/r02b_ncu_summarize.log 2>&1
ncu -i gpurun_out/r02b_argmax.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_argmax_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_argmax.ncu-rep --page source --csv -k regex:'lift_argmax_row' > gpurun_out/r02b_ncu_source_lift_argmax_row.csv 2>/dev/null
gzip -f gpurun_out/r02b_ncu_source_lift_argmax_row.csv
ls -la gpurun_out/ | grep r02b
tail -n 4 gpurun_out/r02b_sanitizer_*.log gpurun_out/r02b_ncu_argmax.log
cat gpurun_out/r02b_ncu_argmax_kernels.md
