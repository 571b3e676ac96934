"""Whole-pass time under scheduling experiments (env knobs are read once per process):
LDIFF_ARGMAX_SMEM_PAD, LDIFF_SIDE_PRIOS (sampler,lifts,tissue,cell), LDIFF_MAIN_PRIO."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from ldiffusion_b200.pipeline import HotPath, synth_inputs

dev = torch.device("cuda")
B, H, W, K, N = 8, 1024, 1024, 11, 5
sets = [synth_inputs(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800, seed=s) for s in (1, 2)]
hp = HotPath(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800)
st = torch.cuda.Stream(priority=int(os.environ.get("LDIFF_MAIN_PRIO", "0")))
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
    for rep in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(200):
            gs[i & 1].replay()
        e1.record(st)
        st.synchronize()
        res.append(e0.elapsed_time(e1) / 200 * 1e3)
ops.check_status(dev)
print(os.environ.get("TAG", ""), " ".join(f"{r:.1f}" for r in sorted(res)), "us/pass", flush=True)
