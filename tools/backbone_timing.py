"""Time the OUT-OF-SCOPE backbone separately (north_star: "the SD UNet backbone stays a cuDNN
call that is timed separately and outside scope").

diffusers / SD weights are not available offline, so this builds random-init torch modules with
the SD-v1.5 architecture constants: UNet2DCondition — 4->4 channels, block widths
(320, 640, 1280, 1280), 2 ResBlocks per level, 1-layer transformers (self-attn, cross-attn 768,
GEGLU) on the first three down levels / mid / last three up levels, 8 heads, GroupNorm 32, SiLU,
time embedding 1280; AutoencoderKL decoder — 4->3 channels, widths (512, 512, 256, 128), 3
ResBlocks per up level, one attention block in the middle.  Everything is a library call
(cuDNN convolutions, cuBLAS GEMMs, flash SDPA).  Not part of bench.py.

    python tools/backbone_timing.py [--batch 8] [--hw 1024]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F


class Res(nn.Module):
    def __init__(self, cin, cout, temb=None):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(32, cin), nn.Conv2d(cin, cout, 3, 1, 1)
        self.t = nn.Linear(temb, cout) if temb else None
        self.n2, self.c2 = nn.GroupNorm(32, cout), nn.Conv2d(cout, cout, 3, 1, 1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.c1(F.silu(self.n1(x)))
        if self.t is not None:
            h = h + self.t(F.silu(temb))[:, :, None, None]
        h = self.c2(F.silu(self.n2(h)))
        return h + (x if self.skip is None else self.skip(x))


class Attn(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.h = heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim, bias=False), nn.Linear(ctx_dim, dim, bias=False),
                                          nn.Linear(ctx_dim, dim, bias=False), nn.Linear(dim, dim))

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, N, C = x.shape
        q = self.q(x).view(B, N, self.h, C // self.h).transpose(1, 2)
        k = self.k(ctx).view(B, ctx.shape[1], self.h, C // self.h).transpose(1, 2)
        v = self.v(ctx).view(B, ctx.shape[1], self.h, C // self.h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.o(o.transpose(1, 2).reshape(B, N, C))


class Transformer(nn.Module):
    def __init__(self, dim, ctx_dim=768, heads=8):
        super().__init__()
        self.norm, self.pin, self.pout = nn.GroupNorm(32, dim), nn.Conv2d(dim, dim, 1), nn.Conv2d(dim, dim, 1)
        self.n1, self.a1 = nn.LayerNorm(dim), Attn(dim, dim, heads)
        self.n2, self.a2 = nn.LayerNorm(dim), Attn(dim, ctx_dim, heads)
        self.n3, self.ff1, self.ff2 = nn.LayerNorm(dim), nn.Linear(dim, dim * 8), nn.Linear(dim * 4, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.pin(self.norm(x)).flatten(2).transpose(1, 2)
        h = h + self.a1(self.n1(h))
        h = h + self.a2(self.n2(h), ctx)
        a, g = self.ff1(self.n3(h)).chunk(2, dim=-1)
        h = h + self.ff2(a * F.gelu(g))
        return x + self.pout(h.transpose(1, 2).reshape(B, C, H, W))


class UNet(nn.Module):
    def __init__(self, widths=(320, 640, 1280, 1280), temb=1280):
        super().__init__()
        self.temb = nn.Sequential(nn.Linear(widths[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.cin = nn.Conv2d(4, widths[0], 3, 1, 1)
        self.down, chans, c = nn.ModuleList(), [widths[0]], widths[0]
        for i, w in enumerate(widths):
            for _ in range(2):
                self.down.append(nn.ModuleList([Res(c, w, temb), Transformer(w) if i < 3 else None]))
                c = w
                chans.append(c)
            if i < 3:
                self.down.append(nn.ModuleList([nn.Conv2d(c, c, 3, 2, 1), None]))
                chans.append(c)
        self.mid = nn.ModuleList([Res(c, c, temb), Transformer(c), Res(c, c, temb)])
        self.up = nn.ModuleList()
        for i, w in reversed(list(enumerate(widths))):
            for j in range(3):
                self.up.append(nn.ModuleList([Res(c + chans.pop(), w, temb), Transformer(w) if i < 3 else None,
                                              nn.Conv2d(w, w, 3, 1, 1) if (j == 2 and i > 0) else None]))
                c = w
        self.nout, self.cout = nn.GroupNorm(32, c), nn.Conv2d(c, 4, 3, 1, 1)
        self.w0 = widths[0]

    def forward(self, x, t, ctx):
        half = self.w0 // 2
        f = torch.exp(-math.log(10000) * torch.arange(half, device=x.device, dtype=torch.float32) / half)
        a = torch.as_tensor(t, device=x.device, dtype=torch.float32).reshape(-1, 1) * f[None]
        temb = self.temb(torch.cat([a.cos(), a.sin()], dim=-1).to(x.dtype).expand(x.shape[0], -1))
        h = self.cin(x)
        skips = [h]
        for blk, tr in self.down:
            h = blk(h, temb) if isinstance(blk, Res) else blk(h)
            if tr is not None:
                h = tr(h, ctx)
            skips.append(h)
        h = self.mid[2](self.mid[1](self.mid[0](h, temb), ctx), temb)
        for blk, tr, upc in self.up:
            h = blk(torch.cat([h, skips.pop()], dim=1), temb)
            if tr is not None:
                h = tr(h, ctx)
            if upc is not None:
                h = upc(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.cout(F.silu(self.nout(h)))


class VAEDecoder(nn.Module):
    def __init__(self, widths=(512, 512, 256, 128)):
        super().__init__()
        self.cin = nn.Conv2d(4, widths[0], 3, 1, 1)
        self.m1, self.mn, self.ma, self.m2 = (Res(widths[0], widths[0]), nn.GroupNorm(32, widths[0]),
                                              Attn(widths[0], widths[0], 1), Res(widths[0], widths[0]))
        self.up, c = nn.ModuleList(), widths[0]
        for i, w in enumerate(widths):
            self.up.append(nn.ModuleList([Res(c, w), Res(w, w), Res(w, w),
                                          nn.Conv2d(w, w, 3, 1, 1) if i < len(widths) - 1 else None]))
            c = w
        self.nout, self.cout = nn.GroupNorm(32, c), nn.Conv2d(c, 3, 3, 1, 1)

    def forward(self, z):
        h = self.m1(self.cin(z))
        B, C, H, W = h.shape
        h = h + self.ma(self.mn(h).flatten(2).transpose(1, 2)).transpose(1, 2).reshape(B, C, H, W)
        h = self.m2(h)
        for r1, r2, r3, upc in self.up:
            h = r3(r2(r1(h)))
            if upc is not None:
                h = upc(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.cout(F.silu(self.nout(h)))


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=1024)
    args = ap.parse_args()
    dev, dt = "cuda", torch.bfloat16
    torch.manual_seed(0)
    unet, dec = UNet().to(dev, dt).eval(), VAEDecoder().to(dev, dt).eval()
    B, h = args.batch, args.hw // 8
    z = torch.randn(B, 4, h, h, device=dev, dtype=dt)
    ctx = torch.randn(B, 77, 768, device=dev, dtype=dt)
    ms_unet = timeit(lambda: unet(z, 501, ctx))
    ms_vae = timeit(lambda: dec(z))
    out = {"unet_ms_per_call": ms_unet, "vae_decode_ms_per_call": ms_vae, "batch": B, "hw": args.hw, "dtype": "bf16",
           "unet_params_M": sum(p.numel() for p in unet.parameters()) / 1e6,
           "vae_decoder_params_M": sum(p.numel() for p in dec.parameters()) / 1e6,
           "per_5_step_batch_ms": 5 * (ms_unet + ms_vae),
           "note": "random-init SD-v1.5-shaped modules, eager PyTorch library calls; out of scope, context only"}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/backbone_timing.json", "w"), indent=1)


if __name__ == "__main__":
    main()
