"""Single-GPU cost of the fused exchange inside the pass (W=1 self-push) vs the plain pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from ldiffusion_b200.dist import ConfusionExchange
from ldiffusion_b200.pipeline import HotPath, synth_inputs

dev = torch.device("cuda")
B, H, W, K, N = 8, 1024, 1024, 11, 5
sets = [synth_inputs(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800, seed=s) for s in (1, 2)]


def bench(mode):
    hp = HotPath(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800)
    if mode != "plain":
        x = ConfusionExchange(K, 2, _local_group=(0, 1))
        ConfusionExchange.connect_local([x])
        hp.attach_exchange(x, deferred=(mode == "deferred"))
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
        best = 1e9
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(200):
                gs[i & 1].replay()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
    ops.check_status(dev)
    return best


for mode in ("plain", "immediate", "deferred", "plain", "immediate", "deferred", "plain", "deferred"):
    print(f"{mode:10s} {bench(mode):8.2f} us/pass", flush=True)

