"""Whole-pass time at the bench shape (8 x 1024^2, K=11, 5 steps, bf16) for a list of configurations, in one
process: each configuration = (label, fused, decode-tail staging variant, lift+argmax variant, chain priorities).

    python tools/pass_time.py                 # default sweep
    python tools/pass_time.py "('x', True, 2, 0, None)" ...
"""
import ast
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import _cabi, ops
from ldiffusion_b200.pipeline import HotPath, synth_inputs

dev = torch.device("cuda")
B, H, W, K, N = 8, 1024, 1024, int(os.environ.get("PASS_K", "11")), 5
sets = [synth_inputs(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800, seed=s) for s in (1, 2)]
lib = _cabi.lib()

DEFAULT = [
    ("unfused, reg-staged tails, per-pixel argmax (round 1)", False, 0, 4, None),
    ("unfused, TMA tails", False, 1, 4, None),
    ("fused, reg-staged tails", True, 0, 0, None),
    ("fused, TMA 4x2", True, 1, 0, None),
    ("fused, TMA 3x3", True, 2, 0, None),
    ("fused, TMA 2x4", True, 3, 0, None),
    ("fused, TMA 2x3", True, 4, 0, None),
    ("fused, TMA 4x2, per-pixel argmax", True, 1, 4, None),
]
configs = [ast.literal_eval(a) for a in sys.argv[1:]] or DEFAULT
ref = None
for cfg in configs:
    label, fused, tma, amax, prios = cfg[:5]
    dstreams = cfg[5] if len(cfg) > 5 else 1
    skip = cfg[6] if len(cfg) > 6 else ()                  # chains left out (where does the pass's time go?)
    lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, tma)
    lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, amax)
    hp = HotPath(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800)
    hp.fused = hp.fused and fused
    hp.decode_streams = dstreams
    if len(cfg) > 7:
        for k, v in cfg[7].items():                          # extra HotPath attributes, e.g. {'tissue_hist_fused': False}
            setattr(hp, k, v)
    for name in skip:
        setattr(hp, "_chain_" + name, lambda *a, **k: None)
    if prios is not None:
        hp.CHAIN_PRIORITIES = tuple(prios)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for s in sets:
            hp.run(s)
        gs = []
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                hp.run(s)
            gs.append(g)
        for i in range(20):
            gs[i & 1].replay()
        st.synchronize()
        res = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(200):
                gs[i & 1].replay()
            e1.record(st)
            st.synchronize()
            res.append(e0.elapsed_time(e1) / 200 * 1e3)
    ops.check_status(dev)
    out = {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in hp.results().items()}
    same = ""
    if skip:
        pass
    elif ref is None:
        ref = out
    else:
        same = "  results_equal=" + str(all(
            all(torch.equal(x, y) for x, y in zip(a if isinstance(a, list) else [a], b if isinstance(b, list) else [b]))
            for a, b in ((ref[k], out[k]) for k in ref)))
    print(f"{label:58s} launches={hp.launches_per_pass():2d}  " + " ".join(f"{r:.1f}" for r in sorted(res)) + " us/pass" + same,
          flush=True)
lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, 1)
lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, 0)
