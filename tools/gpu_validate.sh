#!/bin/bash
# whole-state GPU validation in one gpurun call (~4 GPU-minutes): GPU suite, smoke, bench line, per-kernel bench, pass
# sweeps, ncu launch list of the bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -m pytest tests -q -m gpu ) > gpurun_out/val_gpu_tests.log 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/val_smoke.log 2>&1
python bench.py > gpurun_out/val_bench_1gpu.json 2> gpurun_out/val_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_bench_reference.json 2>> gpurun_out/val_bench_1gpu.err
python tools/kbench_fused.py > gpurun_out/val_kbench_fused.txt 2>&1
python tools/kbench.py > gpurun_out/val_kbench.txt 2>&1
python tools/pass_time.py > gpurun_out/val_pass_time.txt 2>&1
( echo "== decode tails 2 stages x 2 CTAs/SM (library default)"; python tools/pass_overlap.py; echo "== decode tails 2 stages x 1 CTA/SM (what HotPathRing selects from three passes in flight)"; NFLY=1,2,3,4 TMA=11 python tools/pass_overlap.py ) > gpurun_out/val_pass_overlap.txt 2>&1
P='(-1,-1,0,0,-2)'
python tools/pass_time.py "('all chains', True, 6, 0, $P)" "('decode tails only', True, 6, 0, $P, 1, ('sampler','lifts','tissue','cell'))" "('all but the decode tails', True, 6, 0, $P, 1, ('decode',))" "('tissue chain only', True, 6, 0, $P, 1, ('decode','sampler','lifts','cell'))" "('cell chain only', True, 6, 0, $P, 1, ('decode','sampler','lifts','tissue'))" "('lift chain only', True, 6, 0, $P, 1, ('decode','sampler','tissue','cell'))" "('sampler chain only', True, 6, 0, $P, 1, ('decode','lifts','tissue','cell'))" > gpurun_out/val_pass_breakdown.txt 2>&1
LDIFF_BENCH_EXTRAS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/val_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/val_ncu_bench.log 2>&1
for f in gpurun_out/val_gpu_tests.log gpurun_out/val_smoke.log; do tail -n 4 $f; done
cut -c1-300 gpurun_out/val_bench_1gpu.json
cat gpurun_out/val_pass_overlap.txt gpurun_out/val_pass_breakdown.txt
wc -l gpurun_out/val_launches_bench.csv
