#!/bin/bash
# eight GPUs of one box: the bench under torchrun (confusion sum once per evaluation, the default)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/g8_bench_8gpu.json 2> gpurun_out/g8_bench_8gpu.err
tail -n 3 gpurun_out/g8_bench_8gpu.err
cut -c1-300 gpurun_out/g8_bench_8gpu.json
LDIFF_BENCH_EXTRAS=0 timeout 200 python bench.py > gpurun_out/g8_bench_1gpu_same_box.json 2>> gpurun_out/g8_bench_8gpu.err
cut -c1-300 gpurun_out/g8_bench_1gpu_same_box.json
