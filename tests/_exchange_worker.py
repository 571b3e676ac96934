"""torchrun worker of tests/test_gpu_exchange.py::test_two_processes_over_cuda_ipc (one rank per GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    from ldiffusion_b200 import ops
    from ldiffusion_b200.dist import ConfusionExchange, init_from_env
    rank, world = init_from_env("nccl")
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    K, n = 11, 1 << 21
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    x = ConfusionExchange(K, channels=2, device=dev)
    C = torch.zeros(2, K + 1, K, dtype=torch.int64, device=dev)
    out = torch.zeros_like(C)
    preds = [torch.randint(0, K, (n,), device=dev, dtype=torch.uint8, generator=gen) for _ in range(4)]
    gts = [torch.randint(0, K + 2, (n,), device=dev, dtype=torch.uint8, generator=gen) for _ in range(4)]
    ops.status_word(dev)
    s = torch.cuda.Stream(dev)
    graphs = []
    with torch.cuda.stream(s):
        for k in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                C.zero_()
                x.hist_push(preds[2 * k], gts[2 * k], C[0], channel=0)
                x.hist_push(preds[2 * k + 1], gts[2 * k + 1], C[1], channel=1)
                x.reduce(out=out)
            graphs.append(g)
        for step in range(12):
            graphs[step & 1].replay()
            s.synchronize()
            ref = C.clone()
            dist.all_reduce(ref)
            assert torch.equal(out, ref), f"rank {rank} step {step}: exchange != NCCL all-reduce"
            assert int(out.sum()) == 2 * n * world
    # once-per-evaluation sum through the stand-alone push, against NCCL
    total = torch.randint(0, 1 << 40, (2, K + 1, K), device=dev, dtype=torch.int64, generator=gen)
    ref = total.clone()
    dist.all_reduce(ref)
    assert torch.equal(x.allreduce(total), ref), f"rank {rank}: push + reduce != NCCL all-reduce"
    # per-image metrics need the per-tile matrices: one all_gather in global tile order (evaluate.py:95-102)
    from ldiffusion_b200.dist import gather_tile_confusions, shard_tiles
    n_tiles = 7
    mine = shard_tiles(n_tiles, rank, world)
    local = torch.stack([torch.full((K + 1, K), 1000 * t + 1, dtype=torch.int64, device=dev) for t in mine])
    allm = gather_tile_confusions(local, n_tiles)
    assert allm.shape == (n_tiles, K + 1, K)
    for t in range(n_tiles):
        assert int(allm[t, 0, 0]) == 1000 * t + 1 and int(allm[t].min()) == int(allm[t].max())
    ops.check_status(dev)
    dist.barrier()
    x.close()
    print("EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
