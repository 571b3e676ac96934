#!/usr/bin/env python
"""bench.py — throughput of the L-Diffusion sampling-and-feature hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the whole hot path (ldiffusion_b200/pipeline.py) over
one batch of 8 synthetic 1024x1024 PUMA-shaped patches per GPU, K=11 classes,
5 sampling steps, bf16 storage / fp32 arithmetic (BASELINE.json configs[1]).
The SD-v1.5 UNet / VAE are cuDNN library calls outside the scope (north_star):
their outputs are synthetic tensors resident in HBM.  Prints ONE JSON line.

* value      whole-job patches/s with the inputs resident in HBM, the pass replayed
             as a CUDA graph (two rotating input sets, ~0.7 GB > L2)
* e2e        the same metric through the public API (HotPath.run_host) from pinned
             HOST buffers: every step copies all inputs host->device and all results
             back; copy-in / compute / copy-out are pipelined over three streams
* roofline   the dominant kernel (decode_tail_gray) timed alone with CUDA events
* cpu_baseline  the oracle's restatement of the reference's PyTorch CPU op chains
             on this box's host cores (bounded sample), N=1 only
* --impl reference  times that CPU path as the reference arm
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1024x1024 patches/sec through the sampling-and-feature hot path (5 steps)"
UNIT = "patches/s"
WORKLOAD = "cell_1024x1024_b8_k11_5step"
B, H, W, K, NSTEPS, NINST = 8, 1024, 1024, 11, 5, 800


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1, help="patches per CPU-baseline step")
    return ap.parse_args()


AR_NOTE = {"peer": "histogram kernel pushes into NVLink peer windows + one-block reduce, inside the graph",
           "nccl": "NCCL all-reduce behind each graph"}


def config(n_gpus):
    return {"workload": WORKLOAD, "batch_per_gpu": B, "patch": [H, W], "classes": K, "sampling_steps": NSTEPS,
            "instances_per_patch": NINST, "storage": "bf16", "arithmetic": "fp32",
            "parallelism": "single GPU" if n_gpus == 1 else
                           f"patch-sharded x{n_gpus} (no data-path collective; the two int64 confusion matrices are "
                           f"summed across ranks every step: {AR_NOTE.get(os.environ.get('LDIFF_ALLREDUCE', 'peer'), '')})",
            "l2": "two rotating input sets of 0.35 GB each (> 126 MB L2)",
            "backbone": "SD-v1.5 UNet/VAE outputs are synthetic resident tensors (cuDNN calls, out of scope)"}


# ----------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's eager op chains on host cores
# ----------------------------------------------------------------------------

def cpu_chain_setup(n_patches):
    import torch
    from ldiffusion_b200.pipeline import synth_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    inp = synth_inputs(n_patches, H, W, K, NSTEPS, dtype=torch.float32, device="cpu", n_instances=NINST, seed=1234)
    g = torch.Generator().manual_seed(1234)
    hw = torch.randn(K, 256, generator=g) / 16
    cw = torch.randn(K, 256, generator=g) / 16
    zb = torch.zeros(K)
    return inp, (hw, zb, cw, zb)


def cpu_chain_step(inp, weights):
    from oracle.pipeline import run_chain      # the oracle, timed as the CPU baseline
    return run_chain(inp, K, *weights)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n = args.cpu_sample
    inp, weights = cpu_chain_setup(n)
    for _ in range(args.warmup):
        cpu_chain_step(inp, weights)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_chain_step(inp, weights)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"{n} patch(es) of {WORKLOAD} per step, fp32, all stages incl. per-image metric chains"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python and cannot travel to the GPU box; this is the oracle's restatement "
                "of its eager PyTorch op chains (oracle/pipeline.py) on the host cores",
    }))


# ----------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------

class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (t0 - 0.15 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from ldiffusion_b200 import _cabi, ops
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()
    from ldiffusion_b200.dist import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local) if world > 1 else "single process: not bound"
    torch.set_num_threads(max(1, min(8, len(os.sched_getaffinity(0)))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dt = torch.bfloat16
    # ---- inputs: pinned host copies (for e2e) and two device-resident sets (for value)
    host = synth_inputs(B, H, W, K, NSTEPS, dtype=dt, device="cpu", n_instances=NINST, seed=1234 + rank, pin=True)
    dev_sets = []
    for s in range(2):
        hs = host if s == 0 else synth_inputs(B, H, W, K, NSTEPS, dtype=dt, device="cpu", n_instances=NINST,
                                              seed=4321 + rank)
        dev_sets.append(HotPathInputs(*[([t.to(dev) for t in f] if isinstance(f, list) else f.to(dev))
                                        for f in (hs.latents, hs.eps, hs.decoded, hs.head_feat, hs.inst_map,
                                                  hs.inst_feats, hs.gt)]))
    hp = HotPath(B, H, W, K, NSTEPS, dtype=dt, device=dev, n_instances=NINST, seed=1234 + rank)
    # N > 1: the only cross-rank step is the sum of the two int64 confusion matrices.  Default "peer":
    # the histogram kernels push their matrices into every rank's window over NVLink peer memory and a
    # one-block kernel adds the W rows, all inside the pass (graph-captured, no collective library on
    # the path).  LDIFF_ALLREDUCE=nccl keeps the separate NCCL all-reduce behind each pass instead.
    ar_mode = os.environ.get("LDIFF_ALLREDUCE", "peer") if world > 1 else "none"
    if ar_mode == "peer":
        from ldiffusion_b200.dist import ConfusionExchange
        hp.attach_exchange(ConfusionExchange(K, channels=2, device=dev),
                           deferred=os.environ.get("LDIFF_XCHG_DEFERRED", "1") == "1")
    launches_per_pass = hp.launches_per_pass()

    # ---- value: graph-replayed passes over resident inputs (N > 1: the exchange is part of the graph;
    # in nccl mode an eager async all-reduce follows each graph, LDIFF_GRAPH_ALLREDUCE=1 captures it).
    stream = torch.cuda.Stream(dev)
    nccl_ar = ar_mode == "nccl"
    graph_ar = nccl_ar and os.environ.get("LDIFF_GRAPH_ALLREDUCE", "0") == "1"
    graphs = []
    barrier()                                         # ranks enter the first exchanged pass together
    with torch.cuda.stream(stream):
        for s in range(2):
            hp.run(dev_sets[s])                       # warm-up outside capture
            if nccl_ar:
                dist.all_reduce(hp.C)                 # creates the NCCL communicator
        stream.synchronize()
        c0 = _cabi.launch_count()
        for s in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                hp.run(dev_sets[s])
                if graph_ar:
                    dist.all_reduce(hp.C)
            graphs.append(g)
        assert (_cabi.launch_count() - c0) == 2 * launches_per_pass, "launch count claim is wrong"

    # the all-reduce of step i runs on NCCL's stream while the graph of step i+1 already executes:
    # the matrices are staged into one of two small buffers so the next pass may zero hp.C
    Cred = [torch.empty_like(hp.C) for _ in range(2)]
    works = [None, None]

    def step(i):
        graphs[i & 1].replay()
        if nccl_ar and not graph_ar:
            k = i & 1
            if works[k] is not None:
                works[k].wait()                       # stream-side wait, the host does not block
            Cred[k].copy_(hp.C)
            works[k] = dist.all_reduce(Cred[k], async_op=True)

    def drain():
        for k in range(2):
            if works[k] is not None:
                works[k].wait()
                works[k] = None

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.1)
    t_load0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
        drain()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        ev0.record(stream)
        for i in range(args.steps):
            step(i)
        drain()
        ev1.record(stream)
        barrier()
        t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    ops.check_status(dev)
    if ar_mode == "peer":                             # untimed check: the in-pass exchange equals an NCCL all-reduce
        with torch.cuda.stream(stream):
            ref = hp.C.clone()
            dist.all_reduce(ref)
            assert torch.equal(ref, hp.flush_exchange()), "peer exchange differs from the NCCL all-reduce"
            assert int(hp.C_global.sum()) == 2 * world * B * H * W, "exchange lost pixels"

    # ---- roofline of the dominant kernel: decode_tail_gray alone, rotating over 10 decoded tensors (0.5 GB)
    imgs = dev_sets[0].decoded + dev_sets[1].decoded
    reps = 4 * len(imgs)
    with torch.cuda.stream(stream):
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk, stream=stream):
            for r in range(reps):
                ops.decode_tail_gray(imgs[r % len(imgs)], want_rgb=False, gray_out=hp.planes[:, r % NSTEPS])
        gk.replay()
        stream.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(5):
            gk.replay()
        k1.record(stream)
        stream.synchronize()
    kernel_us = k0.elapsed_time(k1) / (5 * reps) * 1e3
    alg_bytes = B * H * W * (3 * 2 + 1)               # 3 bf16 planes in, 1 gray byte out, per pixel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kernel_us * 1e-6) / 1e9
    traffic = None
    try:                                              # DRAM bytes per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["decode_tail_vec16_kernel<bf16,gray>"]
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    # compulsory HBM bytes of one whole pass: what the pass would cost at the copy roofline
    px = B * H * W
    pass_bytes = (NSTEPS * px * 7 + px * 3            # decode tails (+ RGB on the last step)
                  + px * 2                            # label plane copy
                  + (NSTEPS + 1) * B * 3 * 64 * 64 * 2 * 32 + B * 3 * H * W * 2   # lifts: 2x2 footprints (sectors), RGB up
                  + px * 5                            # LUT paint
                  + 2 * px * 2                        # two confusion passes
                  + B * 256 * 32 * 32 * 2 + px        # head features in, mask out
                  + B * NINST * 256 * 2               # instance features
                  + NSTEPS * (2 + 4) * B * 4 * 128 * 128 * 2)   # sampler (approx. 6 latent tensors per step)

    # ---- e2e: pinned host inputs -> H2D -> pass -> D2H of every result, through the public API
    # (HotPath.run_host: copy-in / compute / copy-out pipelined over three streams)
    h2d = host.nbytes()
    host_out = [hp.alloc_host_results() for _ in range(2)]
    d2h = sum(t.numel() * t.element_size() for t in host_out[0].values())
    after = (lambda: dist.all_reduce(hp.C)) if nccl_ar else None
    e2e_steps = max(4, min(args.steps, 32))          # long enough that the pipeline's fill and drain are < 2 % of it
    with torch.cuda.stream(stream):
        hp.run_host([host] * 3, host_out, after_run=after)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hp.run_host([host] * e2e_steps, host_out, after_run=after)
        e1.record(stream)
        barrier()
    e2e_ms = e0.elapsed_time(e1)
    ops.check_status(dev)
    t_load1 = time.perf_counter()
    if rank == 0:
        time.sleep(0.05)
        clocks.stop()

    # ---- max over ranks
    times = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": launches_per_pass * args.steps, "host_binding": numa,
            "roofline": {"bound": "hbm", "kernel": "decode_tail_vec16_kernel<bf16> (gray)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "kernel_us": kernel_us, "algorithmic_bytes_per_launch": alg_bytes,
                         "how": "kernel alone, CUDA-graph of back-to-back launches over 10 rotating "
                                "[8,3,1024,1024] bf16 inputs (0.5 GB), CUDA events on the launching stream"},
            # nvidia-smi cannot sample faster than ~20 ms and the timed region is short, so the window
            # is every GPU-busy phase of this run (warm-up, timed steps, roofline probe, e2e steps)
            "pass_roofline": {"algorithmic_bytes_per_pass": pass_bytes,
                              "achieved_gbs": pass_bytes / (ms / args.steps * 1e-3) / 1e9,
                              "frac_of_hbm_peak": pass_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                              "note": "whole pass incl. the ALU-bound lift+argmax and launch-bound sampler kernels"},
            "clocks": dict(clocks.summary(t_load0, t_load1), window="warm-up .. end of e2e",
                           timed_region_ms=(t_wall1 - t_wall0) * 1e3),
        }
        if world == 1:
            n = args.cpu_sample
            inp, weights = cpu_chain_setup(n)
            cpu_chain_step(inp, weights)
            t0 = time.perf_counter()
            reps_cpu = 0
            while reps_cpu < 3 or (time.perf_counter() - t0 < 10 and reps_cpu < 20):
                cpu_chain_step(inp, weights)
                reps_cpu += 1
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n * reps_cpu / dtc, "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": f"{reps_cpu} x {n} patch(es) of {WORKLOAD}, fp32, oracle/pipeline.py"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
