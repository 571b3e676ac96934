"""End to end (pinned host slab -> H2D -> pass -> D2H) with the label image as int32 and as Cellpose's uint16, same
box, alternating: what the 2 B/pixel less over the host link are worth (HotPath.run_host, the bench's e2e method)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from ldiffusion_b200.pipeline import HotPath, synth_inputs

dev = torch.device("cuda")
B, H, W, K, N, STEPS = 8, 1024, 1024, 11, 5, 32
main = torch.cuda.current_stream()
runs = {}
for dt in (torch.int32, torch.uint16):
    host = synth_inputs(B, H, W, K, N, dtype=torch.bfloat16, device="cpu", n_instances=800, seed=1, inst_dtype=dt).packed()
    hp = HotPath(B, H, W, K, N, dtype=torch.bfloat16, device=dev, n_instances=800)
    runs[dt] = (host, hp, [hp.alloc_host_results() for _ in range(2)])
for rep in range(3):
    for dt, (host, hp, out) in runs.items():
        hp.run_host([host] * 3, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        hp.run_host([host] * STEPS, out)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / STEPS
        ops.check_status(dev)
        print(f"rep {rep} {str(dt):12s} h2d {hp.host_bytes_per_step(host)[0] / 1e6:7.1f} MB  {ms:6.3f} ms/step  "
              f"{B / ms * 1e3:7.1f} patches/s", flush=True)
a, b = runs[torch.int32][2][0], runs[torch.uint16][2][0]
print("results identical:", all(torch.equal(a[k], b[k]) for k in a if k not in ("_slab",)))
