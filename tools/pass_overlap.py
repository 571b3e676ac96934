"""Throughput of the pass when W passes are in flight at once: W HotPath instances (own buffers, own side streams),
one captured graph each, replayed round-robin on W streams.  W = 1 is tools/pass_time.py's number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import _cabi, ops
from ldiffusion_b200.pipeline import HotPath, synth_inputs

dev = torch.device("cuda")
B, H, W_, K, N = 8, 1024, 1024, 11, 5
lib = _cabi.lib()
if "TMA" in os.environ:                                   # decode-tail pipeline shape (default: the library's)
    lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, int(os.environ["TMA"]))
if "AMAX" in os.environ:                                  # lift+argmax kernel (0 = row form where it applies, 1 = column form)
    lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, int(os.environ["AMAX"]))
for nfly in [int(x) for x in os.environ.get("NFLY", "1,2,3").split(",")]:
    sets = [synth_inputs(B, H, W_, K, N, dtype=torch.bfloat16, device=dev, n_instances=800, seed=s) for s in range(1, nfly + 1)]
    if nfly == 1:
        sets.append(synth_inputs(B, H, W_, K, N, dtype=torch.bfloat16, device=dev, n_instances=800, seed=9))
    hps = [HotPath(B, H, W_, K, N, dtype=torch.bfloat16, device=dev, n_instances=800) for _ in range(max(nfly, 2) if nfly == 1 else nfly)]
    if nfly == 1:
        hps = [hps[0], hps[0]]
    for hp in hps:
        hp.tissue_hist_fused = os.environ.get("TF", "0") == "1"
    streams = [torch.cuda.Stream() for _ in range(nfly)]
    graphs = []
    for i, (hp, s) in enumerate(zip(hps, sets)):
        st = streams[i % nfly]
        with torch.cuda.stream(st):
            hp.run(s)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                hp.run(s)
        graphs.append((g, st))
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    res = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _, st in graphs:
            st.wait_stream(main)
        e0.record(main)
        for _, st in graphs:
            st.wait_stream(main)
        n = 200
        for i in range(n):
            g, st = graphs[i % len(graphs)]
            with torch.cuda.stream(st):
                g.replay()
        for _, st in graphs:
            main.wait_stream(st)
        e1.record(main)
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / n * 1e3)
    ops.check_status(dev)
    print(f"passes in flight = {nfly}: " + " ".join(f"{r:.1f}" for r in sorted(res)) + " us/pass", flush=True)
