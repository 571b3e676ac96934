"""GPU parity of one whole pass of the hot path (HotPath.run) against the oracle's
CPU pipeline, eager and as a replayed CUDA graph, plus the pipelined host API."""
import numpy as np
import pytest
import torch

from oracle import bilinear as obil
from oracle import head as ohead
from oracle import laplace as olap
from oracle import metrics as omet
from oracle.pipeline import run_chain

pytestmark = pytest.mark.gpu

CFG = dict(batch=2, height=256, width=256, num_classes=11, num_steps=5, n_instances=40)


def _setup(dtype, seed=7):
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    host = synth_inputs(CFG["batch"], CFG["height"], CFG["width"], CFG["num_classes"], CFG["num_steps"],
                        dtype=dtype, device="cpu", head_hw=(8, 8), n_instances=CFG["n_instances"], seed=seed)
    dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                          for f in (host.latents, host.eps, host.decoded, host.head_feat, host.inst_map,
                                    host.inst_feats, host.gt)])
    hp = HotPath(dtype=dtype, device="cuda", head_hw=(8, 8), seed=seed, **CFG)
    return host, dev, hp


def _check(host, hp, dtype):
    K, n = CFG["num_classes"], CFG["num_steps"]
    res = {k: (v.cpu() if torch.is_tensor(v) else [t.cpu() for t in v]) for k, v in hp.results().items()}
    ref = run_chain(host, K, hp.head_w.cpu(), hp.head_b.cpu(), hp.cell_w.cpu(), hp.cell_b.cpu())
    exact = dtype == torch.float32
    # a-2: latents after the last step
    if exact:
        assert torch.equal(res["latents"], ref["lat"][-1])
    else:
        torch.testing.assert_close(res["latents"].float(), ref["lat"][-1], rtol=2e-2, atol=2e-2)
    # a-1: the kernel's own Philox stream, restated on the CPU
    blocks = (host.latents.numel() + 3) // 4
    for i, t in enumerate(hp.scheduler._host_timesteps):
        nz = olap.laplace_philox(host.latents.numel(), hp.scheduler.laplace_scale(t), hp.seed, i * blocks)
        want = (host.latents.float().reshape(-1) + nz).reshape(host.latents.shape)
        torch.testing.assert_close(res["noisy"][i].float(), want, rtol=1e-5 if exact else 1e-2,
                                   atol=1e-6 if exact else 1e-2)
    # a-3: integer exact (given the same decoded tensors, fp32 or bf16)
    assert np.array_equal(res["pixel_planes"].numpy(), ref["pixel_planes"])
    assert np.array_equal(res["rgb"].numpy(), ref["rgb"])
    # a-4
    if exact:
        want = obil.feature_concat_spec([d.numpy() for d in host.decoded])
        assert np.array_equal(res["featcat"].numpy(), want)
        torch.testing.assert_close(res["featcat"], ref["featcat"], rtol=1e-3, atol=1e-6)
        small = obil.lift_spec(host.decoded[-1].numpy(), (64, 64))
        assert np.array_equal(res["rgb_up"].numpy(), obil.lift_spec(small, (CFG["height"], CFG["width"])))
    else:
        torch.testing.assert_close(res["featcat"].float(), ref["featcat"], rtol=2e-2, atol=2e-2)
    assert torch.equal(res["label_small"], ref["label_small"])
    # a-5: logits within tolerance, masks exact given the kernel's own logits
    torch.testing.assert_close(res["logits"], ref["logits"], rtol=1e-3, atol=1e-3)
    assert np.array_equal(res["mask_tissue"].numpy(),
                          ohead.lift_argmax_spec(res["logits"].numpy(), (CFG["height"], CFG["width"])))
    assert (res["mask_cell"] != ref["mask_cell"]).float().mean() < 2e-3   # only if an instance logit flips
    # a-6: exact given the masks
    want_c = np.stack([omet.confusion_matrix(res[m].numpy(), host.gt.numpy(), K) for m in ("mask_tissue", "mask_cell")])
    assert np.array_equal(res["confusion"].numpy(), want_c)
    if exact and torch.equal(res["mask_tissue"], ref["mask_tissue"]) and torch.equal(res["mask_cell"], ref["mask_cell"]):
        assert np.array_equal(res["confusion"].numpy(), ref["confusion"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pass_matches_oracle_pipeline(dtype):
    from ldiffusion_b200 import _cabi, ops
    host, dev, hp = _setup(dtype)
    c0 = _cabi.launch_count()
    hp.run(dev)
    assert _cabi.launch_count() - c0 == hp.launches_per_pass()
    torch.cuda.synchronize()
    ops.check_status("cuda")
    _check(host, hp, dtype)


def test_graph_replay_equals_eager():
    host, dev, hp = _setup(torch.bfloat16, seed=11)
    hp.run(dev)
    torch.cuda.synchronize()
    eager = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in hp.results().items()}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            hp.run(dev)
    for k, v in hp.results().items():                       # scribble over the outputs, then replay
        for t in (v if isinstance(v, list) else [v]):
            t.zero_()
    g.replay()
    torch.cuda.synchronize()
    for k, v in hp.results().items():
        for a, b in zip(v if isinstance(v, list) else [v], eager[k] if isinstance(eager[k], list) else [eager[k]]):
            assert torch.equal(a, b), k


def test_run_host_pipelined_matches_device_pass():
    from ldiffusion_b200.pipeline import synth_inputs
    host, dev, hp = _setup(torch.bfloat16, seed=13)
    hosts = [synth_inputs(CFG["batch"], CFG["height"], CFG["width"], CFG["num_classes"], CFG["num_steps"],
                          dtype=torch.bfloat16, device="cpu", head_hw=(8, 8), n_instances=CFG["n_instances"],
                          seed=20 + i, pin=True) for i in range(3)]
    outs = [hp.alloc_host_results() for _ in range(3)]
    hp.run_host(hosts, outs)
    torch.cuda.synchronize()
    for hb, out in zip(hosts, outs):
        from ldiffusion_b200.pipeline import HotPathInputs
        d = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                            for f in (hb.latents, hb.eps, hb.decoded, hb.head_feat, hb.inst_map, hb.inst_feats, hb.gt)])
        hp.run(d)
        torch.cuda.synchronize()
        res = hp.results()
        for k in hp.RESULT_KEYS:
            assert torch.equal(out[k], res[k].cpu()), k


# ---- BASELINE.json configs as parity cases ---------------------------------------------------------

@pytest.mark.parametrize("K", [6, 7])
def test_tissue_config_classes(K):
    """configs[2]: tissue level, K=6 (and the reference's own K=7), through the whole pass."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    cfg = dict(CFG, num_classes=K)
    host = synth_inputs(cfg["batch"], cfg["height"], cfg["width"], K, cfg["num_steps"], dtype=torch.bfloat16,
                        device="cpu", head_hw=(8, 8), n_instances=cfg["n_instances"], seed=3)
    dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                          for f in (host.latents, host.eps, host.decoded, host.head_feat, host.inst_map,
                                    host.inst_feats, host.gt)])
    hp = HotPath(dtype=torch.bfloat16, device="cuda", head_hw=(8, 8), seed=3, **cfg)
    hp.run(dev)
    torch.cuda.synchronize()
    ops.check_status("cuda")
    res = hp.results()
    assert int(res["mask_tissue"].max()) < K and int(res["mask_cell"].max()) < K
    assert np.array_equal(res["mask_tissue"].cpu().numpy(),
                          ohead.lift_argmax_spec(res["logits"].cpu().numpy(), (cfg["height"], cfg["width"])))
    want = np.stack([omet.confusion_matrix(res[m].cpu().numpy(), host.gt.numpy(), K) for m in ("mask_tissue", "mask_cell")])
    assert np.array_equal(res["confusion"].cpu().numpy(), want)


def test_tiling_64_tiles_sharded_matches_single_pass():
    """configs[3]: 64 tiles, tile i -> rank i mod W for W in 1/2/4/8 (ranks emulated one after the
    other on this GPU): the summed per-rank matrices equal the single-pass matrix and the oracle,
    bit for bit, and the per-tile matrices give the reference's per-image averages."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.dist import shard_tiles
    from ldiffusion_b200.metrics import summarize_images
    K, T, S = 11, 64, 128
    rng = np.random.default_rng(0)
    preds = rng.integers(0, K, (T, S, S)).astype(np.uint8)
    gts = rng.integers(0, K + 1, (T, S, S)).astype(np.uint8)
    gts[gts == K] = 255
    pd, gd = torch.from_numpy(preds).cuda(), torch.from_numpy(gts).cuda()
    single = ops.confusion_hist(pd.view(-1), gd.view(-1), K).cpu().numpy()
    assert np.array_equal(single, omet.confusion_matrix(preds, gts, K))
    per_tile = ops.confusion_hist_batched(pd, gd, K).cpu().numpy()
    assert np.array_equal(per_tile.sum(0), single)
    for W in (1, 2, 4, 8):
        total = np.zeros_like(single)
        for r in range(W):
            idx = shard_tiles(T, r, W)
            total += ops.confusion_hist(pd[idx].contiguous().view(-1), gd[idx].contiguous().view(-1), K).cpu().numpy()
        assert np.array_equal(total, single), W
    want = omet.evaluate_images_chain(list(preds[:6]), list(gts[:6]), K)
    got = summarize_images(per_tile[:6])
    for k in want:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k


def test_sampler_stress_config():
    """configs[4]: [32,4,64,64], set_timesteps(50) -> 51 fused step launches + Laplace noising,
    eager and graph-captured: latents bit-exact against the oracle."""
    from ldiffusion_b200 import LaplacePLMSScheduler
    from oracle.scheduler import PNDMOracle
    g = torch.Generator().manual_seed(50)
    shape = (32, 4, 64, 64)
    x0 = torch.randn(shape, generator=g) * 5.5
    ref = PNDMOracle(); ref.set_timesteps(50)
    eps = [torch.randn(shape, generator=g) for _ in ref.timesteps]
    x = x0
    for e, t in zip(eps, ref.timesteps):
        x = ref.step(e, t, x)
    sch = LaplacePLMSScheduler()
    eps_d = [e.cuda() for e in eps]
    bufs = [torch.empty(shape, device="cuda") for _ in eps]
    noisy = torch.empty(shape, device="cuda")

    def loop(xd):
        sch.set_timesteps(50)
        for i, t in enumerate(sch._host_timesteps):
            sch.add_laplace_noise(xd, t, seed=1, offset=i)          # fused noise launch (result unused here)
            xd = sch.step(eps_d[i], t, xd, out=bufs[i]).prev_sample
        return xd

    assert torch.equal(loop(x0.cuda()).cpu(), x)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    xs = x0.cuda()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            out = loop(xs)
    bufs[-1].zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), x)


def test_pass_full_size_patches():
    """BASELINE configs[1] geometry (1024x1024, K=11, 5 steps, bf16, 800 instances) on 2 patches:
    every integer output of the pass equals the oracle pipeline; masks equal the pinned decision
    rule on the kernel's own logits."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    B, H, W, K, n = 2, 1024, 1024, 11, 5
    host = synth_inputs(B, H, W, K, n, dtype=torch.bfloat16, device="cpu", seed=99)
    dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                          for f in (host.latents, host.eps, host.decoded, host.head_feat, host.inst_map,
                                    host.inst_feats, host.gt)])
    hp = HotPath(B, H, W, K, n, dtype=torch.bfloat16, device="cuda", seed=99)
    hp.run(dev)
    torch.cuda.synchronize()
    ops.check_status("cuda")
    res = hp.results()
    ref = run_chain(host, K, hp.head_w.cpu(), hp.head_b.cpu(), hp.cell_w.cpu(), hp.cell_b.cpu(), metrics="none")
    assert np.array_equal(res["pixel_planes"].cpu().numpy(), ref["pixel_planes"])
    assert np.array_equal(res["rgb"].cpu().numpy(), ref["rgb"])
    assert torch.equal(res["label_small"].cpu(), ref["label_small"])
    torch.testing.assert_close(res["logits"].cpu(), ref["logits"], rtol=1e-3, atol=1e-3)
    mt = res["mask_tissue"].cpu().numpy()
    assert np.array_equal(mt, ohead.lift_argmax_spec(res["logits"].cpu().numpy(), (H, W)))
    assert (mt != ref["mask_tissue"].numpy()).mean() < 1e-4              # logits differ by summation order only
    assert (res["mask_cell"].cpu() != ref["mask_cell"]).float().mean() < 2e-3
    want = np.stack([omet.confusion_matrix(res[m].cpu().numpy(), host.gt.numpy(), K) for m in ("mask_tissue", "mask_cell")])
    assert np.array_equal(res["confusion"].cpu().numpy(), want)
    assert int(res["confusion"].sum()) == 2 * B * H * W
