"""Multi-GPU plumbing: tiles shard across ranks, the confusion matrix is the
only thing exchanged.

One process per GPU (torchrun).  Stages a-1..a-5 are per-patch with no
cross-patch state, so tile ``i`` simply goes to rank ``i mod W`` and no data-path
collective is needed.  Evaluation adds the integer matrices: one
``all_reduce(SUM)`` of ``(K+1)*K`` int64 (1056 bytes for K=11, latency-bound,
order-independent and therefore bit-exact at any world size), issued straight on
the buffer the histogram kernel accumulated into.  Per-image metrics (the
reference averages per-image Dice, ``evaluate.py:95-102``) need the per-tile
matrices: one ``all_gather`` of ``[tiles_per_rank, K+1, K]``.

The reference's closest analogue is nnU-Net's ``all_gather_object`` of pickled
tp/fp/fn arrays (``nnUNetTrainer.py:1004-1012``).
"""
import os
from typing import List

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None):
    """Initialise torch.distributed from torchrun's environment (no-op single process)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
    return world()


def shard_tiles(n_tiles: int, rank: int = None, world_size: int = None) -> List[int]:
    """Tile i -> rank i mod W."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, n_tiles, world_size))


def allreduce_confusion(C: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of an int64 confusion matrix (no-op single process)."""
    if C.dtype != torch.int64:
        raise TypeError("confusion matrices are int64")
    _, ws = world()
    if ws > 1:
        dist.all_reduce(C, op=dist.ReduceOp.SUM)
    return C


def gather_tile_confusions(local: torch.Tensor, n_tiles: int) -> torch.Tensor:
    """local: int64 [tiles_on_this_rank, K+1, K] in shard_tiles order -> int64
    [n_tiles, K+1, K] in global tile order on every rank (one all_gather)."""
    rank, ws = world()
    if ws == 1:
        return local
    per_rank = (n_tiles + ws - 1) // ws
    pad = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((ws,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1, *pad.shape[1:]), pad) if buf.is_cuda else \
        dist.all_gather(list(buf.unbind(0)), pad)
    out = torch.empty((n_tiles,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(ws):
        idx = list(range(r, n_tiles, ws))
        out[idx] = buf[r, : len(idx)]
    return out


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPUs local to
    its GPU's PCIe root.  With one process per GPU every rank streams ~0.3 GB per step over its
    own link; without this the buffers of several ranks can land on one socket and the
    host-to-device copies then queue on the inter-socket link.  Returns a short description."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return "no local cpus in the allowed set"
        os.sched_setaffinity(0, cpus)
        return f"{bdf} -> cpus {cpulist}"
    except Exception as e:                       # sysfs layout / permissions differ between hosts
        return f"not bound ({type(e).__name__})"
