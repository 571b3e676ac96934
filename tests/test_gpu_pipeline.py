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


def _mask_report(ours, ref_mask, ref_scores, our_scores, what):
    """Masks against the literal chain: every differing pixel / instance must be one where the REFERENCE's own
    top-2 score gap is below twice the measured contraction difference (the only way two correct evaluations of the
    same dot products can order two classes differently).  Returns the mismatch count."""
    diff = ours != ref_mask
    n = int(diff.sum())
    if n:
        delta = float((our_scores - ref_scores).abs().max())
        top2 = torch.topk(ref_scores, 2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1])[diff]
        assert float(gap.max()) <= 2 * delta + 1e-7, (what, n, float(gap.max()), delta)
    return n


def _check(host, hp, dtype, cfg=None):
    cfg = cfg or CFG
    K, n, H, W = cfg["num_classes"], cfg["num_steps"], cfg["height"], cfg["width"]
    res = {k: (v.cpu() if torch.is_tensor(v) else [t.cpu() for t in v]) for k, v in hp.results().items()}
    exact = dtype == torch.float32
    # the oracle pass with the build's storage type emulated (fp32 arithmetic, one rounding per stored tensor)
    # and the pinned-rounding lifts: every floating output below is compared BIT FOR BIT in both dtypes
    ref = run_chain(host, K, hp.head_w.cpu(), hp.head_b.cpu(), hp.cell_w.cpu(), hp.cell_b.cpu(), feat_size=hp.feat_size,
                    storage=None if exact else dtype, spec_lifts=True)
    # a-2: latents after the last step
    assert torch.equal(res["latents"].float(), ref["lat"][-1])
    # a-1: the kernel's own Philox stream, restated on the CPU; bf16: the restated variate to one storage ulp
    blocks = (host.latents.numel() + 3) // 4
    for i, t in enumerate(hp.scheduler._host_timesteps):
        nz = olap.laplace_philox(host.latents.numel(), hp.scheduler.laplace_scale(t), hp.seed, i * blocks,
                                 storage="f32" if exact else "bf16")
        want = (host.latents.float().reshape(-1) + nz).reshape(host.latents.shape)
        if exact:
            torch.testing.assert_close(res["noisy"][i], want, rtol=1e-5, atol=1e-6)      # device log vs libm: <= 6e-6
        else:
            err = (res["noisy"][i].float() - want).abs()
            # one bf16 rounding of the sum + the hardware log2's error in the noise (<= 6e-6 relative, 1.7e-7 b absolute)
            assert (err <= want.abs() * 2.0 ** -8 + nz.abs().reshape(want.shape) * 2e-5 + 3e-7).all()
            assert (res["noisy"][i].float() != want.to(dtype).float()).float().mean() < 2e-3
    # a-3: integer exact (given the same decoded tensors, fp32 or bf16)
    assert np.array_equal(res["pixel_planes"].numpy(), ref["pixel_planes"])
    assert np.array_equal(res["rgb"].numpy(), ref["rgb"])
    # a-4: bit-exact against the spec tier in the storage dtype (and within the 1e-3 contract of the ATen chain)
    assert torch.equal(res["featcat"].float(), ref["featcat"])
    assert torch.equal(res["rgb_up"].float(), ref["rgb_up"])
    if exact:
        torch.testing.assert_close(res["featcat"], obil.feature_concat_chain([d for d in host.decoded], hp.feat_size),
                                   rtol=1e-3, atol=1e-6)
    assert torch.equal(res["label_small"], ref["label_small"])
    # a-5: logits within the 1e-3 contract (summation order is the library's / the tensor core's)
    torch.testing.assert_close(res["logits"], ref["logits"], rtol=1e-3, atol=1e-3)
    #      tissue mask: bit-exact decision rule on the kernel's own logits ...
    assert np.array_equal(res["mask_tissue"].numpy(), ohead.lift_argmax_spec(res["logits"].numpy(), (H, W)))
    #      ... and against the literal chain every differing pixel is a reference near-tie
    from oracle.bilinear import lift_chain
    n_t = _mask_report(res["mask_tissue"], ref["mask_tissue"], lift_chain(ref["logits"], (H, W)),
                       lift_chain(res["logits"], (H, W)), "tissue")
    #      cell mask: the painted class of an instance changes only if its reference logits are a near-tie
    n_c = int((res["mask_cell"] != ref["mask_cell"]).sum())
    if n_c:
        from oracle.head import cell_classify_chain
        ours_l = hp_cell_logits(hp, host)
        for b in range(host.inst_feats.shape[0]):
            _, ref_l = cell_classify_chain(host.inst_feats[b].float(), hp.cell_w.cpu().float(), hp.cell_b.cpu().float())
            _mask_report(ours_l[b][:, 1:].argmax(1), ref_l[:, 1:].argmax(1), ref_l[:, 1:], ours_l[b][:, 1:], "cell")
    # a-6: exact given the masks — ALWAYS; and against the oracle's matrices it differs by exactly the moved pixels
    want_c = np.stack([omet.confusion_matrix(res[m].numpy(), host.gt.numpy(), K) for m in ("mask_tissue", "mask_cell")])
    assert np.array_equal(res["confusion"].numpy(), want_c)
    moved = np.abs(res["confusion"].numpy() - ref["confusion"]).sum(axis=(1, 2))
    assert moved[0] <= 2 * n_t and moved[1] <= 2 * n_c
    if n_t == 0 and n_c == 0:
        assert np.array_equal(res["confusion"].numpy(), ref["confusion"])
    return n_t, n_c


def hp_cell_logits(hp, host):
    """The pass's own instance logits (the kernel can emit them next to the LUT)."""
    from ldiffusion_b200 import ops
    lo = torch.empty(host.inst_feats.shape[:2] + (hp.K,), dtype=torch.float32, device="cuda")
    ops._cell_classify(host.inst_feats.cuda(), hp.cell_w, hp.cell_b, hp.inst_ids, hp.lut.clone(), lo, hp.status)
    return lo.cpu()


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
    """Packed batches (one pinned slab in, one result slab out per step) and plain per-tensor batches."""
    from ldiffusion_b200.pipeline import HotPathInputs, synth_inputs
    host, dev, hp = _setup(torch.bfloat16, seed=13)
    hosts = [synth_inputs(CFG["batch"], CFG["height"], CFG["width"], CFG["num_classes"], CFG["num_steps"],
                          dtype=torch.bfloat16, device="cpu", head_hw=(8, 8), n_instances=CFG["n_instances"],
                          seed=20 + i, pin=True) for i in range(3)]
    for batches in ([h.packed(pin=True) for h in hosts], hosts):
        outs = [hp.alloc_host_results() for _ in range(3)]
        hp.run_host(batches, outs)
        torch.cuda.synchronize()
        h2d, d2h = hp.host_bytes_per_step(batches[0])
        assert h2d == hosts[0].nbytes() and d2h == sum(outs[0][k].numel() * outs[0][k].element_size() for k in hp.RESULT_KEYS)
        for hb, out in zip(hosts, outs):
            d = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda()) for f in hb.fields()])
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


def test_pass_with_uint16_instance_maps_equals_the_int32_pass():
    """The same batch with Cellpose's uint16 label image: every result of the pass is identical, from device
    tensors and through the packed host slab (run_host), and the slab is 2 B/pixel smaller."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    res, nbytes = {}, {}
    for dt in (torch.int32, torch.uint16):
        host = synth_inputs(CFG["batch"], CFG["height"], CFG["width"], CFG["num_classes"], CFG["num_steps"],
                            dtype=torch.bfloat16, device="cpu", head_hw=(8, 8), n_instances=CFG["n_instances"], seed=5,
                            inst_dtype=dt)
        assert host.inst_map.dtype == dt
        dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda()) for f in host.fields()])
        hp = HotPath(dtype=torch.bfloat16, device="cuda", head_hw=(8, 8), seed=5, **CFG)
        hp.run(dev)
        torch.cuda.synchronize()
        ops.check_status("cuda")
        res[dt] = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in hp.results().items()}
        packed = host.packed()
        out = [hp.alloc_host_results()]
        hp.run_host([packed], out)
        torch.cuda.synchronize()
        ops.check_status("cuda")
        assert torch.equal(out[0]["mask_cell"], res[dt]["mask_cell"].cpu())
        assert torch.equal(out[0]["confusion"], res[dt]["confusion"].cpu())
        nbytes[dt] = hp.host_bytes_per_step(packed)[0]
    for k, v in res[torch.int32].items():
        for a, b in zip(v if isinstance(v, list) else [v], res[torch.uint16][k] if isinstance(v, list) else [res[torch.uint16][k]]):
            assert torch.equal(a, b), k
    assert nbytes[torch.int32] - nbytes[torch.uint16] == 2 * CFG["batch"] * CFG["height"] * CFG["width"]


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


@pytest.mark.parametrize("geom", ["configs[1] 1024^2 K=11", "configs[2] 1024^2 K=6 B=8", "configs[0] 1x512^2 K=11"])
def test_pass_at_baseline_config_geometries(geom):
    """BASELINE configs 0 / 1 / 2 at their own geometry (bf16, 5 steps; configs[1] on 2 of its 8 patches to keep the
    CPU oracle short): every stored output bit-exact against the storage-emulating oracle, masks exact on the kernel's
    own logits and reference-near-ties against the literal chain, confusion matrices exact."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    cfg = {"configs[1] 1024^2 K=11": dict(batch=2, height=1024, width=1024, num_classes=11, num_steps=5, n_instances=800, head=(32, 32)),
           "configs[2] 1024^2 K=6 B=8": dict(batch=8, height=1024, width=1024, num_classes=6, num_steps=5, n_instances=800, head=(32, 32)),
           "configs[0] 1x512^2 K=11": dict(batch=1, height=512, width=512, num_classes=11, num_steps=5, n_instances=200, head=(16, 16))}[geom]
    head = cfg.pop("head")
    B, H, W, K, n = cfg["batch"], cfg["height"], cfg["width"], cfg["num_classes"], cfg["num_steps"]
    host = synth_inputs(B, H, W, K, n, dtype=torch.bfloat16, device="cpu", head_hw=head, n_instances=cfg["n_instances"], seed=99)
    dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda()) for f in host.fields()])
    hp = HotPath(dtype=torch.bfloat16, device="cuda", seed=99, head_hw=head, feat_size=(H // 16, W // 16), **cfg)
    assert hp.fused
    hp.run(dev)
    torch.cuda.synchronize()
    ops.check_status("cuda")
    n_t, n_c = _check(host, hp, torch.bfloat16, cfg)
    assert n_t <= 1e-4 * B * H * W and n_c <= 2e-3 * B * H * W          # (reported; each one is proven a near-tie in _check)
    assert int(hp.results()["confusion"].sum()) == 2 * B * H * W


def test_eager_gpu_baseline_chain_matches_cpu_oracle():
    """bench.py's eager-PyTorch-on-the-GPU baseline (oracle/eager_gpu.py) computes what the CPU oracle computes:
    integer outputs identical, floating outputs to fp32 library tolerance, masks equal up to contraction-order
    near-ties."""
    from oracle import eager_gpu
    host, dev, hp = _setup(torch.float32, seed=21)
    w = (hp.head_w, hp.head_b, hp.cell_w, hp.cell_b)
    got = eager_gpu.run_chain(dev, CFG["num_classes"], *w, ieee_fp32=True)
    ref = run_chain(host, CFG["num_classes"], *[t.cpu() for t in w], metrics="none")
    assert np.array_equal(got["pixel_planes"].cpu().numpy(), ref["pixel_planes"])
    assert np.array_equal(got["rgb"].cpu().numpy(), ref["rgb"])
    assert torch.equal(got["label_small"].cpu(), ref["label_small"])
    assert torch.equal(got["lat"][-1].cpu(), ref["lat"][-1]) or torch.allclose(got["lat"][-1].cpu(), ref["lat"][-1], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["featcat"].cpu(), ref["featcat"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(got["logits"].cpu(), ref["logits"], rtol=1e-4, atol=1e-4)
    assert (got["mask_tissue"].cpu() != ref["mask_tissue"]).float().mean() < 1e-3
    assert (got["mask_cell"].cpu() != ref["mask_cell"]).float().mean() < 5e-3
    K = CFG["num_classes"]
    want_c = np.stack([omet.confusion_matrix(got[m].cpu().numpy(), host.gt.numpy(), K) for m in ("mask_tissue", "mask_cell")])
    assert np.array_equal(got["confusion"].cpu().numpy(), want_c)
