"""GPU parity of the fused histogram + peer exchange (include/ldiff.h "a-6 across GPUs"): the sum
every rank reads must equal the sum of the oracle's matrices, bit for bit, at any world size.

Single-GPU cases wire several same-process windows together (each "rank" is a window on the one
GPU); the multi-process case needs two GPUs and is skipped otherwise (run it with
``gpurun --gpus 2``)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import metrics as om

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _group(world, K, channels, timeout_ms=10000):
    from ldiffusion_b200.dist import ConfusionExchange
    g = [ConfusionExchange(K, channels, timeout_ms=timeout_ms, _local_group=(r, world)) for r in range(world)]
    ConfusionExchange.connect_local(g)
    return g


@pytest.mark.parametrize("world,K,channels,lag", [(1, 11, 2, 0), (3, 11, 2, 0), (4, 7, 1, 1), (2, 40, 1, 0)])
def test_push_reduce_equals_sum_of_oracle_matrices(world, K, channels, lag):
    """lag = 1: every reduce is enqueued one step late (after the NEXT step's pushes) and must still
    return the sums of the step it belongs to — pushes and reduces are matched by count."""
    from ldiffusion_b200 import ops
    rng = np.random.default_rng(world * 100 + K)
    g = _group(world, K, channels)
    n = 70001
    C = torch.zeros(world, channels, K + 1, K, dtype=torch.int64, device="cuda")
    want_prev = None
    for step in range(7):                                    # more steps than ring slots
        want = np.zeros((channels, K + 1, K), dtype=np.int64)
        C.zero_()
        for r in range(world):
            for ch in range(channels):
                p = rng.integers(0, K, n, dtype=np.uint8)
                t = rng.integers(0, K + 3, n, dtype=np.uint8)
                want[ch] += om.confusion_matrix(p, t, K)
                g[r].hist_push(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda(), C[r, ch], channel=ch)
        for r in range(world):
            if lag == 0:
                np.testing.assert_array_equal(g[r].reduce().cpu().numpy(), want)
            elif want_prev is not None:
                np.testing.assert_array_equal(g[r].reduce().cpu().numpy(), want_prev)
        want_prev = want
        # the local matrix is still the rank's own (the push does not touch it)
        assert int(C.sum()) == world * channels * n
    ops.check_status("cuda")
    for x in g:
        x.close()


def test_push_reduce_in_cuda_graph():
    from ldiffusion_b200 import ops
    K, world = 11, 2
    g = _group(world, K, 1)
    rng = np.random.default_rng(5)
    p = [torch.from_numpy(rng.integers(0, K, 1 << 20, dtype=np.uint8)).cuda() for _ in range(world)]
    t = [torch.from_numpy(rng.integers(0, K, 1 << 20, dtype=np.uint8)).cuda() for _ in range(world)]
    C = torch.zeros(world, K + 1, K, dtype=torch.int64, device="cuda")
    out = torch.zeros(world, 1, K + 1, K, dtype=torch.int64, device="cuda")
    ops.status_word("cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            C.zero_()
            for r in range(world):
                g[r].hist_push(p[r], t[r], C[r])
            for r in range(world):
                g[r].reduce(out=out[r])
        for _ in range(9):
            graph.replay()
        s.synchronize()
    want = sum(om.confusion_matrix(p[r].cpu().numpy(), t[r].cpu().numpy(), K) for r in range(world))
    for r in range(world):
        np.testing.assert_array_equal(out[r, 0].cpu().numpy(), want)
    ops.check_status("cuda")
    for x in g:
        x.close()


def test_missing_rank_times_out_instead_of_hanging():
    from ldiffusion_b200 import ops
    import time
    K = 5
    g = _group(2, K, 1, timeout_ms=500)
    p = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    C = torch.zeros(K + 1, K, dtype=torch.int64, device="cuda")
    g[0].hist_push(p, p, C)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = g[0].reduce()                                      # rank 1 never pushed
    torch.cuda.synchronize()
    assert 0.4 < time.perf_counter() - t0 < 3.0
    assert int(got[0, 0, 0]) == 4096                         # own row only
    g[0].hist_push(p, p, C)
    t0 = time.perf_counter()
    g[0].reduce()                                            # the timeout is on record: no second wait
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.3
    with pytest.raises(RuntimeError, match="did not deliver"):
        ops.check_status("cuda")
    for x in g:
        x.close()


def test_exchange_argument_errors():
    from ldiffusion_b200 import ops
    from ldiffusion_b200.dist import ConfusionExchange
    g = _group(1, 5, 1)[0]
    p = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        g.hist_push(p, p, torch.zeros(5, 5, dtype=torch.int64, device="cuda"))
    with pytest.raises(ops.LdiffError):
        g.hist_push(p, p, torch.zeros(6, 5, dtype=torch.int64, device="cuda"), channel=3)
    with pytest.raises(ops.LdiffError):
        ConfusionExchange(5, channels=9, _local_group=(0, 1))
    g.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_over_cuda_ipc():
    """One process per GPU, windows mapped through CUDA IPC, graph-captured push + reduce compared
    with an NCCL all-reduce of the same matrices on every step."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "_exchange_worker.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("EXCHANGE_OK") == 2, r.stdout[-2000:]


@pytest.mark.parametrize("mode", ["round-1 launches", "fused pass", "fused pass, fused tissue histogram"])
@pytest.mark.parametrize("deferred,world", [(False, 1), (True, 1), (True, 2)])
def test_pass_with_fused_exchange_matches_plain_pass(deferred, world, mode):
    """HotPath with the exchange attached (every "rank" its own HotPath and window on the one GPU):
    the summed matrices equal the sum of the plain passes' matrices, eagerly and from replayed CUDA
    graphs.  Ranks that share a GPU run one after the other, so only the deferred reduce (which never
    waits for a push that is not enqueued yet) can be exercised with two of them."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, synth_inputs
    B, H, W, K, n = 2, 256, 256, 11, 5
    kw = dict(dtype=torch.bfloat16, device="cuda", n_instances=50)
    inps = [synth_inputs(B, H, W, K, n, seed=10 + r, **kw) for r in range(world)]
    # (the push rides as the tail of whichever kernel finishes a matrix: the stand-alone histogram, the fused
    # paint + histogram, the fused lift + argmax + histogram)
    hkw = dict(kw) if mode == "round-1 launches" else dict(kw, feat_size=(H // 16, W // 16))

    def make():
        hp = HotPath(B, H, W, K, n, seed=7, **hkw)
        assert hp.fused == (mode != "round-1 launches")
        hp.tissue_hist_fused = mode.endswith("tissue histogram")
        return hp

    plain = []
    for r in range(world):
        hp = make()
        hp.run(inps[r])
        plain.append(hp.C.clone())
    want = sum(plain)
    g = _group(world, K, 2)
    hps = [make() for _ in range(world)]
    for r in range(world):
        hps[r].attach_exchange(g[r], deferred=deferred)
        assert hps[r].launches_per_pass() == {"round-1 launches": 31, "fused pass": 18,
                                               "fused pass, fused tissue histogram": 17}[mode]
    for _ in range(3):                                       # eager passes, ranks interleaved
        for r in range(world):
            hps[r].run(inps[r])
    for r in range(world):
        assert torch.equal(hps[r].C, plain[r])
        assert torch.equal(hps[r].flush_exchange(), want)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for r in range(world):
            hps[r].run(inps[r])                              # deferred: leaves one step outstanding, as in steady state
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            for r in range(world):
                hps[r].run(inps[r])
        for r in range(world):
            hps[r].C_global.zero_()
        for _ in range(6):
            graph.replay()
        for r in range(world):
            assert torch.equal(hps[r].flush_exchange(), want)
        s.synchronize()
    ops.check_status("cuda")
    for x in g:
        x.close()


def test_accumulate_over_an_evaluation_and_sum_once():
    """The reference's schedule: matrices add up over the passes of an evaluation on every rank and are
    summed across ranks ONCE at its end — here through the stand-alone push (no histogram kernel to ride on)."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, synth_inputs
    world, B, H, W, K, n = 3, 2, 256, 256, 11, 5
    kw = dict(dtype=torch.bfloat16, device="cuda", n_instances=50)
    g = _group(world, K, 2)
    want = torch.zeros(2, K + 1, K, dtype=torch.int64, device="cuda")
    totals = []
    for r in range(world):
        hp = HotPath(B, H, W, K, n, seed=7, feat_size=(H // 16, W // 16), **kw)
        single = HotPath(B, H, W, K, n, seed=7, feat_size=(H // 16, W // 16), **kw)
        hp.accumulate = True
        hp.reset_confusion()
        for p in range(3):
            inp = synth_inputs(B, H, W, K, n, seed=40 + 10 * r + p, **kw)
            hp.run(inp)
            single.run(inp)
            want += single.C
        totals.append(hp.C.clone())
    assert torch.equal(sum(totals), want) and int(want.sum()) == 2 * world * 3 * B * H * W
    for r in range(world):
        g[r].push(totals[r])
    for r in range(world):
        assert torch.equal(g[r].reduce(), want)
    ops.check_status("cuda")
    for x in g:
        x.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_tiling_config_through_the_exchange(world):
    """configs[3]: 64 tiles, tile i -> rank i mod W; the matrix every rank reads from the
    fused exchange equals the single-pass matrix and the oracle's, bit for bit."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.dist import shard_tiles
    K, T, S = 11, 64, 128
    rng = np.random.default_rng(0)
    preds = rng.integers(0, K, (T, S, S)).astype(np.uint8)
    gts = rng.integers(0, K + 1, (T, S, S)).astype(np.uint8)
    gts[gts == K] = 255
    pd, gd = torch.from_numpy(preds).cuda(), torch.from_numpy(gts).cuda()
    want = om.confusion_matrix(preds, gts, K)
    g = _group(world, K, 1)
    C = torch.zeros(world, K + 1, K, dtype=torch.int64, device="cuda")
    for r in range(world):
        idx = shard_tiles(T, r, world)
        g[r].hist_push(pd[idx].contiguous().view(-1), gd[idx].contiguous().view(-1), C[r])
    for r in range(world):
        np.testing.assert_array_equal(g[r].reduce()[0].cpu().numpy(), want)
    ops.check_status("cuda")
    for x in g:
        x.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """Launch attributes (the 75 KB tcgen05 head, the bulk-TMA decode tails) and the SM count are kept PER DEVICE:
    the same process drives cuda:0 and cuda:1 and gets the oracle's answers on both (a process-wide
    `static bool attr_set` made the head fail on a second device)."""
    from oracle import head as ohead
    from ldiffusion_b200 import ops
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(1, 256, 8, 16, generator=g).bfloat16()
    w = (torch.randn(11, 256, generator=g) / 16).bfloat16()
    img = torch.empty(1, 3, 64, 64).uniform_(-1.2, 1.2, generator=g).bfloat16()
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            mask, logits = ops.head_argmax(feat.to(dev), w.to(dev), None, (256, 512), return_logits=True)
            assert np.array_equal(mask.cpu().numpy(), ohead.lift_argmax_spec(logits.cpu().numpy(), (256, 512)))
            rgb, gray = ops.decode_tail_gray(img.to(dev), want_rgb=True)
            ops.check_status(dev)
            outs.append((logits.cpu(), rgb.cpu(), gray.cpu()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
