"""CPU: host-side logic of the product (scheduler state machine + scalar folding,
metric drop-ins' argument handling, tile sharding + collectives over gloo, world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import metrics as omet
from oracle.scheduler import PNDMOracle


def _emulate_kernel(sample, eps, mode, sc, dA, den):
    """What ldiff_plms_step computes, restated with torch-CPU fp32 ops in the kernel's
    order (mul/add/div each rounded once) — lets the host logic be checked without a GPU."""
    f = lambda v: torch.tensor(v, dtype=torch.float32)                       # noqa: E731
    if mode == 0:
        eh = eps[0]
    elif mode == 1:
        eh = (eps[0] + eps[1]) * f(0.5)
    elif mode == 2:
        eh = (f(3.0) * eps[0] - eps[1]) * f(0.5)
    elif mode == 3:
        eh = ((f(23.0) * eps[0] - f(16.0) * eps[1]) + f(5.0) * eps[2]) / f(12.0)
    else:
        eh = f(1.0 / 24.0) * (((f(55.0) * eps[0] - f(59.0) * eps[1]) + f(37.0) * eps[2]) - f(9.0) * eps[3])
    return f(sc) * sample - (f(dA) * eh) / f(den)


@pytest.mark.parametrize("n_set", [1, 2, 4, 5, 10, 50])
def test_scheduler_host_logic_reproduces_oracle(monkeypatch, n_set):
    from ldiffusion_b200 import ops
    from ldiffusion_b200.scheduler import LaplacePLMSScheduler
    calls = []

    def fake_plms(sample, eps, mode, sc, dA, den, out=None):
        calls.append((mode, len(eps)))
        return _emulate_kernel(sample, list(eps), mode, sc, dA, den)

    monkeypatch.setattr(ops, "plms_step", fake_plms)
    ref, sch = PNDMOracle(), LaplacePLMSScheduler()
    ref.set_timesteps(n_set); sch.set_timesteps(n_set)
    assert sch.timesteps.tolist() == ref.timesteps.tolist() and sch.timesteps.dtype == torch.int64
    assert torch.equal(sch.alphas_cumprod, ref.alphas_cumprod)
    g = torch.Generator().manual_seed(n_set)
    x = xr = torch.randn(64, generator=g) * 5.5
    for t in ref.timesteps:
        eps = torch.randn(64, generator=g)
        xr = ref.step(eps, t, ref.scale_model_input(xr, t))
        x = sch.step(eps, t, sch.scale_model_input(x, t)).prev_sample
        assert torch.equal(x, xr), f"t={int(t)}"
    want_modes = [0, 1, 2, 3] + [4] * 60
    assert [m for m, _ in calls] == want_modes[:len(calls)]
    assert len(sch.ets) == min(4, max(1, len(calls) - 1)) if len(calls) > 1 else len(sch.ets) == 1


def test_scheduler_errors_and_reset():
    from ldiffusion_b200.scheduler import LaplacePLMSScheduler
    s = LaplacePLMSScheduler()
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 1, torch.zeros(1))
    with pytest.raises(ValueError):
        s.set_timesteps(0)
    s.set_timesteps(4)
    assert s.scale_model_input(5, 3) == 5 and s.init_noise_sigma == 1.0 and len(s) == 1000
    assert s.laplace_scale(601) == pytest.approx(0.916615307, rel=2e-7)
    assert s.plan_step(751)[0] == 0


def test_metric_dropins_validate_arguments():
    from ldiffusion_b200 import metrics as pmet
    with pytest.raises(ValueError):
        pmet._pred_labels_u8(torch.zeros(4))
    C = omet.confusion_matrix(np.array([[0, 1, 1]]), np.array([[0, 1, 2]]), 2)
    assert pmet.pa_from_confusion(C)[1] == [1.0, 1.0]
    d, a = pmet.dice_from_confusion(C)
    assert d.dtype == torch.float32 and d.shape == (2,) and a.shape == ()


def test_evaluate_count_mismatch(tmp_path):
    from PIL import Image
    from ldiffusion_b200 import metrics as pmet
    (tmp_path / "p").mkdir(); (tmp_path / "g").mkdir()
    Image.fromarray(np.zeros((4, 4), np.uint8)).save(tmp_path / "p" / "a.png")
    with pytest.raises(ValueError, match="must be equal"):
        pmet.evaluate(str(tmp_path / "p"), str(tmp_path / "g"), 3, str(tmp_path / "o"), device="cpu")


def test_shard_tiles():
    from ldiffusion_b200.dist import shard_tiles
    for n, w in [(64, 1), (64, 2), (64, 8), (10, 4), (3, 8)]:
        shards = [shard_tiles(n, r, w) for r in range(w)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_tiles, K, q):
    import torch.distributed as dist
    from ldiffusion_b200 import dist as ld
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                         # same data on every rank
    preds = rng.integers(0, K, (n_tiles, 24, 24)).astype(np.uint8)
    gts = rng.integers(0, K + 1, (n_tiles, 24, 24)).astype(np.uint8)
    mine = ld.shard_tiles(n_tiles)
    local = torch.from_numpy(np.stack([omet.confusion_matrix(preds[i], gts[i], K) for i in mine])
                             if mine else np.zeros((0, K + 1, K), np.int64))
    total = ld.allreduce_confusion(local.sum(0).clone())
    tiles = ld.gather_tile_confusions(local, n_tiles)
    q.put((rank, total.numpy(), tiles.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [7, 8])
def test_confusion_collectives_world_size_2(n_tiles):
    K, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tiles, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    preds = rng.integers(0, K, (n_tiles, 24, 24)).astype(np.uint8)
    gts = rng.integers(0, K + 1, (n_tiles, 24, 24)).astype(np.uint8)
    per_tile = np.stack([omet.confusion_matrix(preds[i], gts[i], K) for i in range(n_tiles)])
    for _, total, tiles in got:
        assert np.array_equal(total, per_tile.sum(0))          # == single-process matrix, bit for bit
        assert np.array_equal(tiles, per_tile)                  # global tile order restored


def test_contrastive_pair_sampling_rules_and_cpu_refusal():
    """model/loss.py:64-87: 1 % of each class as anchors, positive = another pixel of the class,
    negatives = num_negatives distinct pixels of other classes; classes without enough negatives
    are skipped; the product refuses CPU feature tensors."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.loss import pixel_contrastive_loss, sample_contrastive_pairs
    lab = torch.zeros(2, 1, 32, 32, dtype=torch.uint8)
    lab[0, 0, :16] = 1
    lab[1, 0, 0, :2] = 3
    pb, pa, pq, neg = sample_contrastive_pairs(lab, 400, torch.Generator().manual_seed(1))
    flat = lab.reshape(2, -1)
    for b, a, q, ng in zip(pb.tolist(), pa.tolist(), pq.tolist(), neg):
        assert flat[b, a] == flat[b, q] and a != q
        assert (flat[b, ng.long()] != flat[b, a]).all() and len(set(ng.tolist())) == 400
    # image 0: classes 0 and 1 of 512 px each -> 5 anchors each; image 1: class 3 (2 px, 1022 negatives)
    # gives max(1, 0) = 1 anchor, class 0 (1022 px) has only 2 negatives -> skipped
    assert pb.tolist().count(0) == 10 and pb.tolist().count(1) == 1
    again = sample_contrastive_pairs(lab, 400, torch.Generator().manual_seed(1))
    assert all(torch.equal(x, y) for x, y in zip((pb, pa, pq, neg), again))
    assert sample_contrastive_pairs(torch.zeros(1, 1, 8, 8, dtype=torch.uint8), 16) is None
    with pytest.raises(ops.LdiffError):
        pixel_contrastive_loss(torch.zeros(1, 5, 32, 32), lab[:1])


def _reference_multimodal_loop(rgb, dtm, pipe, controlnet, proj, noise):
    """segmentor.py:322-386 restated op for op on the CPU (per-image loop, B = 1, the
    ``.repeat`` copies, separate 0.18215 passes) with the Laplace(0,1) noise injected."""
    import torch.nn.functional as F
    from oracle.laplace import depth_noising_chain, depth_unnoise_chain
    rgb = F.interpolate(rgb, size=(256, 256), mode="bilinear", align_corners=False)
    dtm = F.interpolate(dtm, size=(256, 256), mode="bilinear", align_corners=False)
    out = []
    for i in range(dtm.shape[0]):
        dtm_i, rgb_i = dtm[i], rgb[i].unsqueeze(0)
        depth_condition = dtm_i.unsqueeze(0).repeat(1, 3, 1, 1)
        latents = pipe.vae.encode(rgb_i).latent_dist.sample() * 0.18215
        depth_resized = F.interpolate(dtm_i.unsqueeze(0), size=(32, 32), mode="bilinear", align_corners=False)
        depth_resized = depth_resized.repeat(1, latents.shape[1], 1, 1)
        latents_noisy, _ = depth_noising_chain(latents, depth_resized, noise=noise[i:i + 1])
        ids = torch.tensor(pipe.tokenizer(["A remote sense image"])["input_ids"], dtype=torch.long)
        text = proj(pipe.text_encoder(ids)["last_hidden_state"].to(torch.float32)).to(torch.float32)
        pipe.scheduler.set_timesteps(1)
        for timestep in pipe.scheduler.timesteps:
            down, mid = controlnet(sample=latents_noisy, timestep=timestep, encoder_hidden_states=text,
                                   controlnet_cond=depth_condition, return_dict=False)
            noise_pred = pipe.unet(latents_noisy, timestep, encoder_hidden_states=text,
                                   down_block_additional_residuals=down, mid_block_additional_residual=mid).sample
        latents_denoised = depth_unnoise_chain(latents_noisy, noise_pred, depth_resized)
        recon = pipe.vae.decode(latents_denoised / 0.18215).sample
        out.append(recon.squeeze(0).permute(1, 2, 0).numpy())
    return out


def test_multimodal_augment_host_logic(monkeypatch):
    """Segmentor.ldiffusion_augment_for_multimodal (batched, broadcast depth map, fused 0.18215
    factors) against the reference's per-image loop, with the three ldiff operators it calls
    emulated by their oracle chains — the GPU parity of those operators is tests/test_gpu_sampler.py."""
    import torch.nn.functional as F
    from oracle.laplace import depth_noising_chain, depth_unnoise_chain
    from ldiffusion_b200 import ops
    from ldiffusion_b200.segmentor import Segmentor
    from ldiffusion_b200.standin import StandInControlNet, StandInPipeline

    def fake_lift(src, size, **kw):
        return F.interpolate(src, size=size, mode="bilinear", align_corners=False)

    def fake_noising(x, scale, *, noise=None, x_mul=1.0, seed=0, **kw):
        assert scale.shape[1] == 1 and noise is not None
        return depth_noising_chain(x, scale, noise=noise, x_mul=x_mul)[0]

    def fake_residual(x, eps, scale, *, out_div=1.0, out=None):
        assert scale.shape[1] == 1
        return depth_unnoise_chain(x, eps, scale, out_div=out_div)

    monkeypatch.setattr(ops, "bilinear_lift", fake_lift)
    monkeypatch.setattr(ops, "laplace_qsample_map", fake_noising)
    monkeypatch.setattr(ops, "scaled_residual", fake_residual)
    torch.manual_seed(0)
    pipe = StandInPipeline("cpu", seed=4)
    controlnet = StandInControlNet().eval()
    seg = object.__new__(Segmentor)                      # the constructor insists on a CUDA device
    seg.device, seg.ldiffusion_proj = torch.device("cpu"), None
    g = torch.Generator().manual_seed(9)
    rgb, dtm = torch.rand(3, 3, 300, 280, generator=g), torch.rand(3, 1, 300, 280, generator=g) * 2
    noise = torch.distributions.Laplace(0.0, 1.0).sample((3, 4, 32, 32))
    with torch.no_grad():
        got = seg.ldiffusion_augment_for_multimodal(rgb, dtm, pipe, pipe.unet, pipe.vae, controlnet, 3, "cpu",
                                                    noise=noise)
        want = _reference_multimodal_loop(rgb, dtm, pipe, controlnet, seg.ldiffusion_proj, noise)
    assert len(got) == 3 and got[0].shape == (256, 256, 3) and got[0].dtype == np.float32
    for a, b in zip(got, want):                          # batched convolutions vs B = 1: not bit-identical
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)


def test_nnunet_compute_metrics_host_logic(monkeypatch):
    """metrics.nnunet_compute_metrics = one confusion histogram + the evaluator's formulas; the histogram is
    emulated by the oracle here (its GPU parity: tests/test_gpu_head_metrics.py)."""
    from ldiffusion_b200 import metrics as pmet
    monkeypatch.setattr(pmet, "confusion_matrix",
                        lambda pred, ref, K: torch.from_numpy(omet.confusion_matrix(pred.numpy(), ref.numpy(), K)))
    rng = np.random.default_rng(4)
    ref = torch.from_numpy(rng.integers(0, 5, (2, 20, 30)).astype(np.uint8))
    pred = torch.from_numpy(rng.integers(0, 4, (2, 20, 30)).astype(np.uint8))
    got = pmet.nnunet_compute_metrics(ref, pred, [1, 2, 3], ignore_label=4)
    keep = ref != 4
    for r in (1, 2, 3):
        tp = int(((ref == r) & (pred == r) & keep).sum()); fp = int(((ref != r) & (pred == r) & keep).sum())
        fn = int(((ref == r) & (pred != r) & keep).sum())
        assert (got[r]["TP"], got[r]["FP"], got[r]["FN"]) == (tp, fp, fn)
        assert got[r]["TN"] == int(keep.sum()) - tp - fp - fn and got[r]["Dice"] == 2 * tp / (2 * tp + fp + fn)


def test_evaluate_command_line_flags(monkeypatch):
    """python -m ldiffusion_b200.evaluate takes the flags of the reference's evaluate.py:129-139, and importing
    the submodule does not break ``ldiffusion_b200.evaluate(...)`` (the module forwards calls to the function)."""
    import importlib
    import ldiffusion_b200 as L
    from ldiffusion_b200 import metrics as pmet
    assert L.evaluate is pmet.evaluate
    ev = importlib.import_module("ldiffusion_b200.evaluate")
    a = ev.parse_args(["--image-dir", "p", "--label-dir", "l", "--num-classes", "7"])
    assert (a.image_dir, a.label_dir, a.num_classes, a.save_dir) == ("p", "l", 7, "./LDiffusion/eval/eval_report")
    with pytest.raises(SystemExit):
        ev.parse_args(["--image-dir", "p"])
    assert ev.evaluate is pmet.evaluate and ev.pixel_accuracy is pmet.pixel_accuracy
    calls = []
    monkeypatch.setattr(ev, "evaluate", lambda *a, **k: calls.append((a, k)) or "report")
    assert L.evaluate("pred", "gt", 7, save_dir="out") == "report"          # L.evaluate is the module now
    assert calls == [(("pred", "gt", 7), {"save_dir": "out"})]


def test_feature_concat_autograd_wiring(monkeypatch):
    """features.feature_concat_autograd: forward = the lift kernel, backward = its adjoint kernel, both emulated
    here by the oracle (chain forward, numpy spec backward) — the gradient that reaches every decoded step must
    be torch's own autograd through ldiffusion.py:240-247."""
    from oracle import bilinear as obil
    from ldiffusion_b200 import features, ops

    def fake_lift(src, size, *, gray=False, out_dtype=None, **kw):
        y = obil.lift_chain(src.float(), size)
        return obil.gray_weighted_chain(y) if gray else y

    def fake_backward(grad_out, src_shape, *, out_channel=0, gray=False):
        return torch.from_numpy(obil.lift_backward_spec(grad_out.numpy(), src_shape, gray=gray))

    monkeypatch.setattr(ops, "bilinear_lift", fake_lift)
    monkeypatch.setattr(ops, "bilinear_lift_backward", fake_backward)
    g = torch.Generator().manual_seed(12)
    steps = [torch.randn(2, 3, 64, 64, generator=g, requires_grad=True) for _ in range(3)]
    feats = features.feature_concat_autograd(steps, (16, 16))
    assert feats.shape == (2, 3, 16, 16) and feats.requires_grad
    assert torch.equal(feats.detach(), obil.feature_concat_chain([s.detach() for s in steps], (16, 16)))
    go = torch.randn(feats.shape, generator=g)
    feats.backward(go)
    want = obil.feature_concat_grad_chain(steps, go, (16, 16))
    for s_, w_ in zip(steps, want):
        assert torch.equal(s_.grad, w_)


def test_warmup_step_host_logic(monkeypatch):
    """LDiffusionModel.warmup_step / train_ldiffusion (ldiffusion.py:121-295 without DeepSpeed) with every
    ldiff operator emulated by its oracle: gradients reach the UNet and the text projection through the
    feature lift and the InfoNCE term, the parameters move, and the loss equals the reference loop's."""
    import types
    from oracle import bilinear as obil
    from oracle.loss import contrastive_loss_chain
    from ldiffusion_b200 import loss as ploss, ops
    from ldiffusion_b200.ldiffusion import LDiffusionModel
    from ldiffusion_b200.standin import StandInPipeline

    def fake_lift(src, size, *, gray=False, out_dtype=None, **kw):
        y = obil.lift_chain(src.float(), tuple(size))
        if src.dtype == torch.uint8:
            return y.to(torch.uint8)
        return obil.gray_weighted_chain(y) if gray else y

    monkeypatch.setattr(ops, "bilinear_lift", fake_lift)
    monkeypatch.setattr(ops, "bilinear_lift_backward", lambda g, shape, *, out_channel=0, gray=False:
                        torch.from_numpy(obil.lift_backward_spec(g.numpy(), shape, gray=gray)))
    monkeypatch.setattr(ops, "laplace_qsample", lambda x, b, *, noise=None, **kw: x + noise)
    seen = {}

    def fake_loss(feats, labels, pairs=None, seed=0, offset=0, **kw):
        seen["feats"], seen["labels"] = feats.detach().clone(), labels.clone()
        pairs = ploss.sample_contrastive_pairs(labels, 64, torch.Generator().manual_seed(offset))
        seen["pairs"] = pairs
        return contrastive_loss_chain(feats, pairs)

    monkeypatch.setattr(ploss, "pixel_contrastive_loss", fake_loss)
    torch.manual_seed(0)
    pipe = StandInPipeline("cpu", seed=5)
    model = object.__new__(LDiffusionModel)                      # the constructor insists on a CUDA device
    model.device, model.linear_layer, model._pipeline_loader, model.diffusion_path = torch.device("cpu"), None, None, "x"
    g = torch.Generator().manual_seed(1)
    image = torch.rand(2, 3, 64, 64, generator=g)
    label = torch.zeros(2, 1, 128, 128, dtype=torch.uint8)
    label[:, :, :64] = 1
    label[:, :, :, 96:] = 2
    proj = torch.nn.Linear(768, 768)
    params = list(pipe.unet.parameters()) + list(proj.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3)
    before = [p.detach().clone() for p in params]
    pipe.scheduler.set_timesteps(2)
    noise = [torch.randn(2, 4, 8, 8, generator=g) for _ in pipe.scheduler.timesteps]
    loss = model.warmup_step(image, label, pipe, pipe.unet, pipe.vae, proj, opt, 2, noise=noise, size=(16, 16))
    assert loss.ndim == 0 and float(loss) > 0 and seen["feats"].shape == (2, 3, 16, 16)
    assert torch.equal(seen["labels"], obil.lift_chain(label.float(), (16, 16)).to(torch.uint8))
    assert float(loss) == pytest.approx(float(contrastive_loss_chain(seen["feats"], seen["pairs"])), rel=1e-6)
    moved = [not torch.equal(a, b.detach()) for a, b in zip(before, params)]
    assert all(moved[:len(list(pipe.unet.parameters()))]) and any(moved[-2:])        # UNet and projection updated
    # the epoch driver: AdamW as configured at ldiffusion.py:167-175, one mean loss per epoch
    monkeypatch.setattr(LDiffusionModel, "warmup_step",
                        lambda self, image, label, *a, **k: torch.tensor(float(k["step_index"] + 1)))
    args = types.SimpleNamespace(num_inference_steps=10, num_epochs=2, diffusion_path="x")
    log = []
    hist = model.train_ldiffusion(args, [(image, None, label)] * 3, pipeline=pipe, log=log)
    assert hist == [2.0, 5.0] and log == [(1, 2.0), (2, 5.0)]
    assert model.linear_layer.in_features == 768 and model.linear_layer.out_features == 768


def test_instance_map_dtypes_and_packed_slab_layout():
    """Host side of the label-image formats: int32 and Cellpose's uint16 (an int16 view counts as unsigned) are
    accepted, anything else is refused before a launch; a packed host slab keeps the map's own dtype (2 B/pixel
    less for uint16) and round-trips its values; the CPU has no painting path (no fallback)."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.pipeline import synth_inputs
    assert ops._ids16(torch.zeros(2, dtype=torch.int32)) is False
    assert ops._ids16(torch.zeros(2, dtype=torch.uint16)) is True and ops._ids16(torch.zeros(2, dtype=torch.int16)) is True
    for dt in (torch.int64, torch.uint8, torch.float32):
        with pytest.raises(TypeError):
            ops._ids16(torch.zeros(2, dtype=dt))
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        ops.lut_paint(torch.zeros(4, 16, dtype=torch.uint16), torch.zeros(8, dtype=torch.uint8))   # CPU tensors
    kw = dict(dtype=torch.bfloat16, device="cpu", head_hw=(2, 2), n_instances=7, seed=2)
    a = synth_inputs(1, 64, 64, 5, 2, **kw)
    b = synth_inputs(1, 64, 64, 5, 2, inst_dtype=torch.uint16, **kw)
    assert a.inst_map.dtype == torch.int32 and b.inst_map.dtype == torch.uint16
    assert np.array_equal(a.inst_map.numpy(), b.inst_map.numpy().astype(np.int32))
    assert a.nbytes() - b.nbytes() == 2 * 64 * 64
    pb = b.packed(pin=False)
    assert pb.inst_map.dtype == torch.uint16 and np.array_equal(pb.inst_map.numpy(), b.inst_map.numpy())
    assert torch.equal(pb.gt, b.gt) and pb.slab.numel() == b.slab_layout()[1]
