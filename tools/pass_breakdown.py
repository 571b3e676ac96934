"""Per-stage GPU time of one HotPath pass at the bench config: each stage captured alone
into a CUDA graph (rotating over two input sets), replayed, timed with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs

B, H, W, K, N = 8, 1024, 1024, 11, 5
dev = torch.device("cuda")
sets = []
for s in range(2):
    hs = synth_inputs(B, H, W, K, N, dtype=torch.bfloat16, device="cpu", seed=100 + s)
    sets.append(HotPathInputs(*[([t.to(dev) for t in f] if isinstance(f, list) else f.to(dev))
                                for f in (hs.latents, hs.eps, hs.decoded, hs.head_feat, hs.inst_map, hs.inst_feats, hs.gt)]))
hp = HotPath(B, H, W, K, N, device=dev)
for s in sets:
    hp.run(s)
torch.cuda.synchronize()
sch = hp.scheduler
ts = sch._host_timesteps
blocks = (hp.lat_elems + 3) // 4


def stage_laplace(inp):
    for i in range(N):
        ops.laplace_qsample(inp.latents, sch.laplace_scale(ts[i]), seed=1, offset=i * blocks, out=hp.noisy[i])


def stage_plms(inp):
    sch.set_timesteps(N - 1)
    x = inp.latents
    for i in range(N):
        x = sch.step(inp.eps[i], ts[i], x, out=hp.lat[i]).prev_sample


def stage_decode(inp):
    for i in range(N):
        ops.decode_tail_gray(inp.decoded[i], want_rgb=False, rgb_out=hp.rgb if i == N - 1 else None,
                             gray_out=hp.planes[:, i])


def stage_down(inp):
    for i in range(N):
        ops.bilinear_lift(inp.decoded[i], hp.feat_size, out=hp.featcat, out_channel=i, gray=True)
    ops.bilinear_lift(inp.gt.unsqueeze(1), hp.feat_size, out=hp.label_small)
    ops.bilinear_lift(inp.decoded[N - 1], hp.feat_size, out=hp.rgb_small)


STAGES = {
    "laplace x5": stage_laplace, "plms x5": stage_plms, "decode_tail x5": stage_decode,
    "lift down x7": stage_down,
    "label copy": lambda inp: ops.copy_planes_u8(inp.gt, hp.planes[:, N]),
    "lift up": lambda inp: ops.bilinear_lift(hp.rgb_small, (H, W), out=hp.rgb_up),
    "head_logits": lambda inp: ops._head_logits(inp.head_feat, hp.head_w, hp.head_b, hp.logits),
    "lift_argmax": lambda inp: ops._lift_argmax(hp.logits, hp.mask_tissue),
    "cell_classify": lambda inp: ops._cell_classify(inp.inst_feats, hp.cell_w, hp.cell_b, hp.inst_ids, hp.lut, None, hp.status),
    "lut_paint": lambda inp: ops.lut_paint(inp.inst_map, hp.lut, out=hp.mask_cell),
    "confusion x2": lambda inp: (ops.confusion_hist(hp.mask_tissue.view(-1), inp.gt.view(-1), K, out=hp.C[0]),
                                 ops.confusion_hist(hp.mask_cell.view(-1), inp.gt.view(-1), K, out=hp.C[1])),
    "whole pass": lambda inp: hp.run(inp),
}
st = torch.cuda.Stream()
tot = 0.0
for name, fn in STAGES.items():
    with torch.cuda.stream(st):
        fn(sets[0]); st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for r in range(10):
                fn(sets[r & 1])
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            g.replay()
        e1.record(st)
        st.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    if name != "whole pass":
        tot += us
    print(f"{name:16s} {us:8.2f} us")
print(f"{'sum of stages':16s} {tot:8.2f} us")

# ---- which chain bounds the concurrent pass?  Re-time the pass with one chain removed.
import types
orig = {n: getattr(HotPath, n) for n in ("_chain_sampler", "_chain_lifts", "_chain_tissue", "_chain_cell")}


def time_pass(label):
    with torch.cuda.stream(st):
        hp.run(sets[0]); st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for r in range(10):
                hp.run(sets[r & 1])
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            g.replay()
        e1.record(st); st.synchronize()
    print(f"{label:28s} {e0.elapsed_time(e1) / 50 * 1e3:8.2f} us")


time_pass("concurrent pass, all chains")
for name in orig:
    setattr(HotPath, name, lambda self, inp: None)
    time_pass("  without " + name)
    setattr(HotPath, name, orig[name])
for name in orig:
    setattr(HotPath, name, lambda self, inp: None)
time_pass("  decode tails + label only")
