"""Multi-GPU plumbing: tiles shard across ranks, the confusion matrix is the
only thing exchanged.

One process per GPU (torchrun).  Stages a-1..a-5 are per-patch with no
cross-patch state, so tile ``i`` simply goes to rank ``i mod W`` and no data-path
collective is needed.  Evaluation adds the integer matrices ((K+1)*K int64, 1056
bytes for K=11; order-independent and therefore bit-exact at any world size):
either ``ConfusionExchange`` — the histogram kernel itself stores its matrix into
every rank's NVLink-mapped window and a one-block kernel adds the rows — or one
NCCL ``all_reduce(SUM)`` issued straight on the buffer the histogram kernel
accumulated into (``allreduce_confusion``).  Per-image metrics (the
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


class ConfusionExchange:
    """The histogram's all-reduce without a collective library: every rank owns a small window in
    device memory that all peers map through CUDA IPC; ``hist_push`` is the confusion histogram
    whose last block stores the finished matrix into every rank's window over NVLink, ``reduce``
    is a one-block kernel that waits for the W rows and adds them (``include/ldiff.h``, "a-6
    across GPUs").  Integer sums, so the result equals ``allreduce_confusion`` bit for bit.

    ``channels`` independent matrices travel per step (the pass has two: tissue and cell).  Every
    rank pushes every channel once per step and reduces once per step; the j-th ``reduce`` returns
    the sum of the j-th pushes, whether it is enqueued right behind them or a step later next to
    the following step's work (at most two steps may be outstanding).  Both kernels can be captured
    into CUDA graphs.

    Replaces nnU-Net's pickled ``all_gather_object`` of tp/fp/fn (``nnUNetTrainer.py:1004-1012``).
    """

    def __init__(self, num_classes: int, channels: int = 1, device=None, timeout_ms: int = 10000, _local_group=None):
        import ctypes
        from . import _cabi
        self.K, self.channels = int(num_classes), int(channels)
        self.n = (self.K + 1) * self.K
        self._lib, self._ct = _cabi.lib(), ctypes
        if _local_group is not None:                       # (rank, world): same-process windows, see connect_local
            self.rank, self.world = _local_group
        else:
            self.rank, self.world = world()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self._lib.ldiff_xchg_create(self.world, self.rank, self.channels, self.n, ctypes.byref(h)))
        self._h = h
        self._check(self._lib.ldiff_xchg_set_timeout(h, int(timeout_ms)))
        if _local_group is None:
            self._connect_ipc()

    def _check(self, rc):
        from . import ops
        ops.check(rc)

    def _connect_ipc(self):
        ct = self._ct
        mine = ct.create_string_buffer(64)
        self._check(self._lib.ldiff_xchg_ipc_handle(self._h, mine))
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine.raw))    # every window is zeroed before its handle exists
        else:
            gathered = [bytes(mine.raw)]
        blob = ct.create_string_buffer(b"".join(gathered), 64 * self.world)
        with torch.cuda.device(self.device):
            self._check(self._lib.ldiff_xchg_connect_ipc(self._h, blob))
        if self.world > 1:
            dist.barrier()                                       # nobody pushes into a window that is not mapped yet

    @staticmethod
    def connect_local(group):
        """Wire several same-process windows to each other (several "ranks" on one GPU: tests)."""
        import ctypes
        arr = (ctypes.c_void_p * len(group))(*[g._h for g in group])
        for g in group:
            g._check(g._lib.ldiff_xchg_connect_local(g._h, arr))

    def hist_push(self, pred: torch.Tensor, gt: torch.Tensor, C: torch.Tensor, channel: int = 0, gt_lut=None):
        """``ops.confusion_hist(pred, gt, K, out=C)`` + push of the finished C to every rank."""
        from . import ops
        if C.dtype != torch.int64 or C.numel() != self.n or not C.is_contiguous():
            raise ValueError("C must be a contiguous int64 [(K+1), K] matrix")
        if pred.dtype != torch.uint8 or gt.dtype != torch.uint8 or pred.numel() != gt.numel():
            raise ValueError("pred and gt must be uint8 maps of the same size")
        ops._cuda(pred, gt, C)
        self._check(self._lib.ldiff_confusion_hist_push(
            pred.data_ptr(), gt.data_ptr(), None if gt_lut is None else gt_lut.data_ptr(), C.data_ptr(),
            pred.numel(), self.K, self._h, int(channel), ops.status_word(pred.device).data_ptr(),
            torch.cuda.current_stream(pred.device).cuda_stream))
        return C

    def push(self, C: torch.Tensor) -> None:
        """Push finished matrices of ALL channels (int64 [channels, K+1, K]) without a histogram kernel to ride
        on — the once-per-evaluation sum (the reference adds its counts at the end of evaluation)."""
        from . import ops
        if C.dtype != torch.int64 or C.numel() != self.channels * self.n or not C.is_contiguous():
            raise ValueError("C must be a contiguous int64 [channels, K+1, K] tensor")
        ops._cuda(C)
        self._check(self._lib.ldiff_xchg_push(self._h, C.data_ptr(), torch.cuda.current_stream(C.device).cuda_stream))

    def allreduce(self, C: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """push + reduce: the sum over ranks of C, through the peer windows."""
        self.push(C)
        return self.reduce(out)

    def reduce(self, out: torch.Tensor = None) -> torch.Tensor:
        """Sum over ranks of the next not-yet-reduced step's matrices: int64 [channels, K+1, K]."""
        from . import ops
        if out is None:
            out = torch.empty(self.channels, self.K + 1, self.K, dtype=torch.int64, device=self.device)
        if out.dtype != torch.int64 or out.numel() != self.channels * self.n or not out.is_contiguous():
            raise ValueError("out must be a contiguous int64 [channels, K+1, K] tensor")
        self._check(self._lib.ldiff_xchg_reduce(self._h, out.data_ptr(), ops.status_word(out.device).data_ptr(),
                                                torch.cuda.current_stream(out.device).cuda_stream))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            torch.cuda.synchronize(self.device)
            self._lib.ldiff_xchg_destroy(self._h)
            self._h = None


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
