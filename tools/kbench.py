"""Per-kernel micro-benchmark at the BASELINE config-2 shapes (dev tool, not the bench contract).

CUDA-event timing over many back-to-back launches on rotating buffers larger than L2."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops

PEAK = 6546.2


def timeit(fn, nbuf, iters=40, warm=3):
    """GPU time per launch: `iters` launches captured into one CUDA graph (no host
    launch gaps), replayed 5 times, CUDA events around the replays."""
    for i in range(warm):
        fn(i % nbuf)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i % nbuf)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * reps) * 1e3     # us


def main():
    dev = "cuda"
    res = {}
    for dt, s in ((torch.float32, 4), (torch.bfloat16, 2)):
        tag = "f32" if s == 4 else "bf16"
        B = 8
        nb = 6
        imgs = [torch.empty(B, 3, 1024, 1024, device=dev, dtype=dt).uniform_(-1.2, 1.2) for _ in range(nb)]
        planes = torch.empty(B, 6, 1024, 1024, dtype=torch.uint8, device=dev)
        rgb = torch.empty(B, 1024, 1024, 3, dtype=torch.uint8, device=dev)
        us = timeit(lambda i: ops.decode_tail_gray(imgs[i], want_rgb=False, gray_out=planes[:, i % 5]), nb)
        byt = B * 1024 * 1024 * (3 * s + 1)
        res[f"decode_tail_gray_{tag}"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
        us = timeit(lambda i: ops.decode_tail_gray(imgs[i], rgb_out=rgb, gray_out=planes[:, i % 5]), nb)
        byt = B * 1024 * 1024 * (3 * s + 4)
        res[f"decode_tail_gray_rgb_{tag}"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
        # bilinear up 64->1024
        src = torch.randn(B, 3, 64, 64, device=dev).to(dt)
        outs = [torch.empty(B, 3, 1024, 1024, device=dev, dtype=dt) for _ in range(4)]
        us = timeit(lambda i: ops.bilinear_lift(src, (1024, 1024), out=outs[i]), 4)
        byt = B * 3 * 1024 * 1024 * s
        res[f"lift_up_{tag}"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
        fc = torch.empty(B, 5, 64, 64, device=dev, dtype=dt)
        us = timeit(lambda i: ops.bilinear_lift(imgs[i], (64, 64), out=fc, out_channel=i % 5, gray=True), nb)
        res[f"lift_down_gray_{tag}"] = (us, 0, 0)
        # sampler at config shape and at a scaled-up shape (roofline probe)
        for name, shape in (("cfg", (8, 4, 128, 128)), ("big", (1024, 4, 128, 128))):
            x = [torch.randn(shape, device=dev).to(dt) for _ in range(3 if name == "big" else 2)]
            e = [torch.randn(shape, device=dev).to(dt) for _ in range(4)]
            o = torch.empty_like(x[0])
            n = x[0].numel()
            us = timeit(lambda i: ops.laplace_qsample(x[i % len(x)], 0.9, seed=1, offset=i, out=o), len(x))
            res[f"laplace_{name}_{tag}"] = (us, 2 * n * s / us / 1e3, 2 * n * s / us / 1e3 / PEAK)
            us = timeit(lambda i: ops.plms_step(x[i % len(x)], e, 4, 1.01, 0.02, 0.9, out=o), len(x))
            res[f"plms4_{name}_{tag}"] = (us, 6 * n * s / us / 1e3, 6 * n * s / us / 1e3 / PEAK)
            # multimodal variant: depth map [B,1,h,w] broadcast over the 4 latent channels
            sc = torch.rand(shape[0], 1, shape[2], shape[3], device=dev).to(dt)
            us = timeit(lambda i: ops.laplace_qsample_map(x[i % len(x)], sc, seed=1, offset=i, x_mul=0.18215, out=o),
                        len(x))
            res[f"laplace_map_{name}_{tag}"] = (us, 2.25 * n * s / us / 1e3, 2.25 * n * s / us / 1e3 / PEAK)
            us = timeit(lambda i: ops.scaled_residual(x[i % len(x)], e[0], sc, out_div=0.18215, out=o), len(x))
            res[f"scaled_residual_{name}_{tag}"] = (us, 3.25 * n * s / us / 1e3, 3.25 * n * s / us / 1e3 / PEAK)
    # head
    feat = torch.randn(8, 256, 32, 32, device=dev).bfloat16()
    w = (torch.randn(11, 256, device=dev) / 16).bfloat16()
    us = timeit(lambda i: ops.head_logits(feat, w, None), 1)
    res["head_logits_bf16"] = (us, 0, 0)
    logits = ops.head_logits(feat, w, None)
    us = timeit(lambda i: ops.lift_argmax(logits, (1024, 1024)), 1)
    res["lift_argmax"] = (us, 8 * 1024 * 1024 / us / 1e3, 8 * 1024 * 1024 / us / 1e3 / PEAK)
    # paint + confusion
    inst = torch.randint(0, 800, (8, 1024, 1024), device=dev, dtype=torch.int32)
    lut = torch.randint(0, 11, (800,), device=dev, dtype=torch.uint8)
    us = timeit(lambda i: ops.lut_paint(inst, lut), 1)
    res["lut_paint"] = (us, 8 * 1024 * 1024 * 5 / us / 1e3, 8 * 1024 * 1024 * 5 / us / 1e3 / PEAK)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_gpu_head_metrics import _blob_labels
    for nimg in (8, 64):
        gt = torch.from_numpy(_blob_labels((nimg, 1024, 1024), 11, 1)).to(dev)
        pr = torch.from_numpy(_blob_labels((nimg, 1024, 1024), 11, 2, 0)).to(dev)
        C = torch.zeros(12, 11, dtype=torch.int64, device=dev)
        us = timeit(lambda i: ops.confusion_hist(pr.view(-1), gt.view(-1), 11, out=C), 1)
        byt = 2 * nimg * 1024 * 1024
        res[f"confusion_blobs_{nimg}"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
    rnd_p = torch.randint(0, 11, (8 * 1024 * 1024,), device=dev, dtype=torch.uint8)
    rnd_g = torch.randint(0, 11, (8 * 1024 * 1024,), device=dev, dtype=torch.uint8)
    C = torch.zeros(12, 11, dtype=torch.int64, device=dev)
    us = timeit(lambda i: ops.confusion_hist(rnd_p, rnd_g, 11, out=C), 1)
    res["confusion_random_8"] = (us, 2 * 8 * 1024 * 1024 / us / 1e3, 2 * 8 * 1024 * 1024 / us / 1e3 / PEAK)
    # widening N2: nnU-Net sliding-window tail (K=7 heads, 512x512 tiles on a 1024x1024 image, fp16)
    from ldiffusion_b200.sliding_window import SlidingWindowAccumulator, tta_merge
    sw = SlidingWindowAccumulator(7, (1024, 1024), (512, 512), True, dev)
    tiles = [(torch.randn(7, 512, 512, device=dev) * 3).half() for _ in range(4)]
    us = timeit(lambda i: sw.add(tiles[i], 256 * (i % 3), 256 * (i % 2)), 4)
    byt = 512 * 512 * (7 * 2 * 3 + 2 * 3)                  # pred in, acc in+out, gauss in, npred in+out
    res["sw_accumulate_tile"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
    us = timeit(lambda i: tta_merge(tiles, [0, 1, 2, 3]), 1)
    byt = 7 * 512 * 512 * 2 * 5
    res["sw_tta_merge_4"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
    sw.acc.copy_((torch.randn(7, 1024, 1024, device=dev) * 30).half())     # sane accumulators (the timing loop above overflowed them)
    sw.npred.fill_(10.0)
    us = timeit(lambda i: sw.finalize(), 1)
    byt = 1024 * 1024 * (7 * 2 + 2 + 1)
    res["sw_finalize_argmax"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
    # widening N3: pixel-contrastive InfoNCE, 8 x 5 x 64 x 64 features, 1024 negatives per anchor
    from ldiffusion_b200.loss import sample_contrastive_pairs
    lab = torch.zeros(8, 1, 64, 64, dtype=torch.uint8)
    for b in range(8):
        for q in range(6):
            lab[b, 0, 8 * q:8 * q + 16, 5 * q:5 * q + 24] = 1 + (b + q) % 6
    pairs = [t.to(dev) for t in sample_contrastive_pairs(lab, 1024, torch.Generator().manual_seed(0))]
    feat = torch.randn(8, 5, 64, 64, device=dev)
    A = pairs[1].numel()
    loss, lse = torch.empty(A, device=dev), torch.empty(A, device=dev)
    grad, gs = torch.zeros_like(feat), torch.full((1,), 1.0 / A, device=dev)
    byt = A * 1026 * 5 * 4                                   # gathered candidate features (L2-resident map)
    us = timeit(lambda i: ops._infonce_forward(feat, *pairs, loss, lse, 0.5), 1)
    res[f"infonce_forward_{A}x1025"] = (us, byt / us / 1e3, byt / us / 1e3 / PEAK)
    us = timeit(lambda i: ops._infonce_backward(feat, *pairs, lse, gs, grad, 0.5), 1)
    res[f"infonce_backward_{A}x1025"] = (us, 2 * byt / us / 1e3, 2 * byt / us / 1e3 / PEAK)
    for k, (us, gbs, fr) in res.items():
        print(f"{k:32s} {us:10.2f} us  {gbs:9.1f} GB/s  {fr:6.3f} of measured peak")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
