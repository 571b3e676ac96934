#!/bin/bash
# two GPUs of one box: the tests that need a second device / a second process, and the bench under torchrun
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_exchange.py -q -m gpu ) > gpurun_out/g2_tests.log 2>&1
tail -n 4 gpurun_out/g2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/g2_bench_2gpu.json 2> gpurun_out/g2_bench_2gpu.err
tail -n 3 gpurun_out/g2_bench_2gpu.err
cut -c1-400 gpurun_out/g2_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/g2_bench_ref_2gpu.json 2>> gpurun_out/g2_bench_2gpu.err
cut -c1-300 gpurun_out/g2_bench_ref_2gpu.json
