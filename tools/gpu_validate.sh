#!/bin/bash
# whole-state GPU validation in one gpurun call (~2.5 GPU-minutes): GPU suite, smoke, bench line, per-kernel bench, ncu launch list of the bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/val_gpu_tests.log 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/val_smoke.log 2>&1
python bench.py > gpurun_out/val_bench_1gpu.json 2> gpurun_out/val_bench_1gpu.err
python tools/kbench.py > gpurun_out/val_kbench.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/val_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/val_ncu_bench.log 2>&1
for f in gpurun_out/val_gpu_tests.log gpurun_out/val_smoke.log; do tail -n 4 $f; done
cut -c1-200 gpurun_out/val_bench_1gpu.json
grep -E "laplace_map|scaled_residual|lift_argmax" gpurun_out/val_kbench.txt
wc -l gpurun_out/val_launches_bench.csv
