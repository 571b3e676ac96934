#!/bin/bash
# A/B of two builds of the library inside ONE gpurun call (box-to-box variance is larger than most kernel changes):
#   tools/ab_pass.sh ab_libs/lib_base.so ab_libs/lib_new.so [rounds]       (env NFLY / TMA as for tools/pass_overlap.py)
cd "$(dirname "$0")/.."
cp ldiffusion_b200/libldiff_sm100.so /tmp/lib_keep.so
for r in $(seq 1 ${3:-2}); do
  for l in "$1" "$2"; do
    cp "$l" ldiffusion_b200/libldiff_sm100.so
    echo "== $l (round $r)"
    python tools/pass_overlap.py
    [ -n "$AB_KBENCH" ] && python tools/kbench_fused.py 2>&1 | grep -E "$AB_KBENCH"
  done
done
cp /tmp/lib_keep.so ldiffusion_b200/libldiff_sm100.so
