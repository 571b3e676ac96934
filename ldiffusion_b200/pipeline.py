"""One pass of the whole hot path over a batch of patches.

``HotPath`` owns every intermediate buffer (so a pass allocates nothing and can
be captured into a CUDA graph) and strings the stages together the way the
reference's loops do, with the SD-v1.5 UNet / VAE calls replaced by the tensors
they would have produced (``eps[i]``, ``decoded[i]``): those are cuDNN library
calls outside the scope of this package and are timed separately.

Stages per batch (reference lines in brackets):

1. for each of the n sampling steps i (``set_timesteps(n-1)``,
   pixel_latent_vector.py:74-83):
   * Laplace forward noising of the clean latents at t_i    [ldiffusion.py:233-237]
   * PLMS reverse update with eps_i                          [segmentor.py:102-104]
   * decode tail of decoded_i -> gray plane i of the pixel
     vectors (+ uint8 RGB on the last step)                  [pixel_latent_vector.py:80-93]
   * bilinear 64x64 + weighted gray -> channel i of the
     training-path feature tensor                            [ldiffusion.py:240-247]
2. label bilinear down + uint8; last decode 64x64 -> HxW     [ldiffusion.py:224-226,251]
3. tissue head: 1x1 conv -> lift -> argmax(softmax) mask     [conductor.py:127,135; segmentor.py:536]
4. cell head: instance classifier -> LUT -> painted mask     [conductor.py:218-231]
5. confusion matrices of both masks vs gt                    [utils.py:55-104; evaluate.py:11-45]
"""
from dataclasses import dataclass
from typing import List, Optional

import os

import torch

from . import _cabi, ops
from .scheduler import LaplacePLMSScheduler


@dataclass
class HotPathInputs:
    """Everything one pass consumes (device tensors)."""
    latents: torch.Tensor            # [B,4,H/8,W/8] VAE-encoded clean latents (storage dtype)
    eps: List[torch.Tensor]          # n x [B,4,H/8,W/8] UNet outputs
    decoded: List[torch.Tensor]      # n x [B,3,H,W] VAE decoder outputs
    head_feat: torch.Tensor          # [B,256,h,w] tissue decoder features
    inst_map: torch.Tensor           # int32 or uint16 (Cellpose's own dtype below 65 536 labels) [B,H,W] instance ids
    inst_feats: torch.Tensor         # [B,N,256] pooled instance features
    gt: torch.Tensor                 # uint8 [B,H,W] ground-truth classes
    slab = None                      # set by packed(): the uint8 tensor all fields are views of

    def tensors(self):
        return [self.latents, *self.eps, *self.decoded, self.head_feat, self.inst_map, self.inst_feats, self.gt]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def fields(self):
        return (self.latents, self.eps, self.decoded, self.head_feat, self.inst_map, self.inst_feats, self.gt)

    # ---- one contiguous slab per batch: a step's host->device transfer is ONE copy instead of 17 -------------
    def slab_layout(self):
        """[(offset, nbytes, shape, dtype)] of every tensor in ``tensors()`` order, 256-byte aligned; total bytes."""
        off, lay = 0, []
        for t in self.tensors():
            nb = t.numel() * t.element_size()
            lay.append((off, nb, tuple(t.shape), t.dtype))
            off += (nb + 255) & ~255
        return lay, off

    def packed(self, pin: bool = True, device=None) -> "HotPathInputs":
        """A copy of this batch whose tensors are views of ONE uint8 slab (``.slab``): pinned host memory by
        default, or device memory (``device=...``) for the receiving side."""
        lay, total = self.slab_layout()
        if device is None:
            slab = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
        else:
            slab = torch.empty(total, dtype=torch.uint8, device=device)
        views = [slab[o:o + nb].view(dt).view(shape) for (o, nb, shape, dt) in lay]
        if device is None:
            for v, t in zip(views, self.tensors()):
                v.copy_(t)
        n = len(self.eps)
        out = HotPathInputs(views[0], views[1:1 + n], views[1 + n:1 + 2 * n], *views[1 + 2 * n:])
        out.slab = slab
        return out


class _nvtx:
    """NVTX range around a chain when LDIFF_NVTX=1 (the reference has no tracing hooks at all,
    SURVEY 5); lets ``ncu --nvtx --nvtx-include`` pick one chain of the pass."""
    on = os.environ.get("LDIFF_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()


class HotPath:
    def __init__(self, batch: int, height: int, width: int, num_classes: int, num_steps: int = 5,
                 dtype=torch.bfloat16, device="cuda", head_hw=(32, 32), feat_size=(64, 64),
                 n_instances: int = 800, head_channels: int = 256, seed: int = 1234):
        dev = torch.device(device)
        self.B, self.H, self.W, self.K, self.n = batch, height, width, num_classes, num_steps
        self.dtype, self.device = dtype, dev
        self.feat_size, self.head_hw, self.n_inst = feat_size, head_hw, n_instances
        self.seed = seed
        self.scheduler = LaplacePLMSScheduler()
        lat = (batch, 4, height // 8, width // 8)
        self.lat_elems = batch * 4 * (height // 8) * (width // 8)
        e = lambda *s, dt=dtype: torch.empty(s, dtype=dt, device=dev)          # noqa: E731
        self.noisy = [e(*lat) for _ in range(num_steps)]
        self.lat = [e(*lat) for _ in range(num_steps - 1)]
        # the results the reference hands to the HOST (RESULT_KEYS) are views of ONE slab, so that a step's
        # device->host transfer is a single copy: final latents, pixel vectors (pixel_latent_vector.py:89-101 writes
        # them to CSV), the uint8 image (a PIL image in the reference), the two masks (np.ndarray at
        # ldiffusion.py:545), the confusion matrices, and the small training-path tensors
        spec = [("latents", lat, dtype), ("pixel_planes", (batch, num_steps + 1, height, width), torch.uint8),
                ("rgb", (batch, height, width, 3), torch.uint8), ("featcat", (batch, num_steps, *feat_size), dtype),
                ("label_small", (batch, 1, *feat_size), torch.uint8),
                ("mask_tissue", (batch, height, width), torch.uint8), ("mask_cell", (batch, height, width), torch.uint8),
                ("confusion", (2, num_classes + 1, num_classes), torch.int64)]
        off, self._out_layout = 0, []
        for name, shape, dt in spec:
            nb = torch.empty(0, dtype=dt).element_size()
            for d in shape:
                nb *= d
            self._out_layout.append((name, off, nb, tuple(shape), dt))
            off += (nb + 255) & ~255
        self.out_slab = torch.zeros(off, dtype=torch.uint8, device=dev)
        ov = {name: self.out_slab[o:o + nb].view(dt).view(shape) for name, o, nb, shape, dt in self._out_layout}
        self.lat.append(ov["latents"])
        self.planes, self.rgb, self.featcat, self.label_small = ov["pixel_planes"], ov["rgb"], ov["featcat"], ov["label_small"]
        self.mask_tissue, self.mask_cell, self.C = ov["mask_tissue"], ov["mask_cell"], ov["confusion"]
        # device-only results (the reference keeps them on the GPU: ldiffusion.py:251-252 feeds rgb_up to the loss)
        self.rgb_small = e(batch, 3, *feat_size)
        self.rgb_up = e(batch, 3, height, width)
        self.logits = e(batch, num_classes, *head_hw, dt=torch.float32)
        self.lut = torch.zeros((batch, n_instances + 1), dtype=torch.uint8, device=dev)
        self.inst_ids = torch.arange(1, n_instances + 1, dtype=torch.int32, device=dev)
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.head_w = (torch.randn(num_classes, head_channels, generator=g) / 16).to(dev, dtype)
        self.head_b = torch.zeros(num_classes, device=dev)
        self.cell_w = (torch.randn(num_classes, head_channels, generator=g) / 16).to(dev, dtype)
        self.cell_b = torch.zeros(num_classes, device=dev)
        self.status = ops.status_word(dev)
        self.exchange = None                                               # set by attach_exchange (multi-GPU)
        # the per-step consumers of the decoder output ride inside the decode tails when the 16x down-sample is
        # the exact 2x2 footprint (H, W multiples of 16 and a feature map of H/16 x W/16); LDIFF_PASS_FUSED=0
        # keeps round 1's separate launches (A/B timing, and the reference point of the equivalence tests)
        self.fused = (os.environ.get("LDIFF_PASS_FUSED", "1") == "1" and height % 16 == 0 and width % 16 == 0
                      and tuple(feat_size) == (height // 16, width // 16) and num_classes <= 15)
        # the tissue chain's histogram: fused into lift+argmax (one launch less, 8 MB less traffic) or its own launch
        # (measured at the bench shape: 118 us per pass with the separate launch against 121 us fused — the histogram's
        # ~8 instructions per pixel land on an issue-bound kernel, while the stand-alone histogram is latency-bound
        # and overlaps the other chains; the fused entry point stays in the library for callers without that overlap)
        self.tissue_hist_fused = os.environ.get("LDIFF_PASS_TISSUE_HIST_FUSED", "0") == "1"
        # accumulate: the confusion matrices are NOT cleared at the start of a pass — they add up over the passes of an
        # evaluation (reset_confusion() starts a new one) and are summed across ranks once at its end, as the
        # reference does; False: every pass starts from zero (and can push its matrices, attach_exchange)
        self.accumulate = False
        self.decode_streams = max(1, min(num_steps, int(os.environ.get("LDIFF_DECODE_STREAMS", "1"))))
        self.decode_tail_shape = None                    # pipeline shape of this pass's decode tails (None: the library's)

    def attach_exchange(self, exchange, deferred: bool = True):
        """Multi-GPU: fuse the cross-rank sum of the two confusion matrices into the pass
        (``dist.ConfusionExchange``, 2 channels).  The histogram kernels push their finished
        matrices into every rank's window over NVLink; the one-block reduce into ``self.C_global``
        runs either at the end of the same pass, or (``deferred``) inside the FOLLOWING pass behind
        its short sampler chain, where it never waits for a slower rank and is off the critical path —
        ``C_global`` then holds the previous pass's sum until ``flush_exchange()``."""
        if exchange.channels != 2 or exchange.K != self.K:
            raise ValueError("the pass exchanges two (K+1) x K matrices")
        self.exchange, self.exchange_deferred, self._unreduced = exchange, bool(deferred), 0
        self.C_global = torch.zeros_like(self.C)

    def reset_confusion(self):
        self.C.zero_()

    def _reduce_pending(self):
        while self.exchange is not None and self._unreduced > 0:
            self.exchange.reduce(out=self.C_global)
            self._unreduced -= 1

    def flush_exchange(self):
        """Reduce the passes whose matrices were pushed but not summed yet (deferred mode) and read the status
        word (a synchronising call: not for use inside a graph capture)."""
        self._reduce_pending()
        if self.exchange is not None:
            ops.check_status(self.device)              # a peer that timed out means C_global is a PARTIAL sum: raise
        return self.C_global

    # number of ldiff kernels one pass launches
    def launches_per_pass(self) -> int:
        xr = 1 if self.exchange is not None else 0                         # the exchange's one-block reduce
        if self.fused:                                                     # sampler n, decode n, lifts 2, tissue 2, cell 2
            return 2 * self.n + 2 + 2 + 2 + xr + (0 if self.tissue_hist_fused else 1)
        return 4 * self.n + 1 + 3 + 2 + 2 + 2 + xr

    def _confusion(self, mask, gt, channel):
        if self.exchange is None:
            ops.confusion_hist(mask.view(-1), gt.view(-1), self.K, out=self.C[channel])
        else:
            self.exchange.hist_push(mask.view(-1), gt.view(-1), self.C[channel], channel=channel)

    def run(self, inp: HotPathInputs, concurrent: bool = True):
        """Enqueue one pass; returns nothing (results live in the preallocated buffers).
        Graph-capturable: no allocation, no sync.

        The pass is five independent chains (sampler, decode tails, up-lift, tissue head, cell head).
        With ``concurrent`` they are forked onto prioritised side streams and joined at the end, so
        inside a CUDA graph the latency-bound latent-sized launches and the classifier chains overlap
        the HBM-bound decode tails instead of queueing behind them (and the caller's stream priority
        does not matter)."""
        cur = torch.cuda.current_stream(self.device)
        n = self.n
        if not self.fused and not self.accumulate:
            self.C.zero_()                                                 # a memset every chain waits for
        if concurrent:
            side = self._side_streams()
            for s in side:
                s.wait_stream(cur)
        else:
            side = [cur] * (4 + self.decode_streams)
        with torch.cuda.stream(side[0]), _nvtx("ldiff.sampler"):
            self._chain_sampler(inp)
            if self.exchange is not None and self.exchange_deferred and self._unreduced > 0:
                # the previous pass's sum rides at the end of the short sampler chain: off the critical
                # path, no extra branch in the graph, and long after every rank has pushed
                self.exchange.reduce(out=self.C_global)
                self._unreduced -= 1
        with torch.cuda.stream(side[1]), _nvtx("ldiff.lifts"):
            self._chain_lifts(inp)
        with torch.cuda.stream(side[2]), _nvtx("ldiff.tissue"):
            self._chain_tissue(inp)
        with torch.cuda.stream(side[3]), _nvtx("ldiff.cell"):
            self._chain_cell(inp)
        for j in range(self.decode_streams):                               # the bandwidth-heavy chain(s)
            with torch.cuda.stream(side[4 + j]), _nvtx("ldiff.decode_tails"):
                self._chain_decode(inp, j)
        if concurrent:
            for s in side:
                cur.wait_stream(s)
        if self.exchange is not None:
            self._unreduced += 1
            if not self.exchange_deferred:
                self._reduce_pending()

    # CUDA stream priorities of the five chains (sampler, lifts, tissue, cell, decode tails); captured
    # graphs keep them per kernel node.  When blocks of several chains are waiting for an SM, the
    # bandwidth-bound decode tails and the latency-bound latent kernels go first and the two
    # register-heavy classifier chains fill in behind them: measured 146 -> 135 us per pass against
    # equal priorities in round 1 (tools/pass_sched.py; giving the classifier chains the high priority
    # instead cost 155 us), and with three passes in flight in the final build 88.6 us against 94.4 us
    # for equal priorities (tools/prio_sweep.sh, LDIFF_SIDE_PRIOS).
    CHAIN_PRIORITIES = (-2, -1, 0, 0, -2)

    def _side_streams(self):
        st = getattr(self, "_side", None)
        if st is None:
            env = os.environ.get("LDIFF_SIDE_PRIOS")
            prios = [int(x) for x in env.split(",")] if env else list(self.CHAIN_PRIORITIES)
            st = [torch.cuda.Stream(self.device, priority=prios[min(i, 4)]) for i in range(4 + self.decode_streams)]
            self._side = st
        return st

    def _chain_sampler(self, inp):
        n, sch = self.n, self.scheduler
        sch.set_timesteps(n - 1)
        ts = sch._host_timesteps
        x = inp.latents
        blocks = (self.lat_elems + 3) // 4
        for i in range(n):
            if self.fused:        # reverse update + forward noising of the step in ONE launch
                x, _ = sch.step_then_noise(inp.eps[i], ts[i], x, inp.latents, seed=self.seed, offset=i * blocks,
                                           out=self.lat[i], noisy_out=self.noisy[i])
            else:
                ops.laplace_qsample(inp.latents, sch.laplace_scale(ts[i]), seed=self.seed, offset=i * blocks,
                                    out=self.noisy[i])
                x = sch.step(inp.eps[i], ts[i], x, out=self.lat[i]).prev_sample

    def _chain_decode(self, inp, part: int = 0):
        """Decode tails of the steps i = part (mod decode_streams).  The n tails have no data dependence on one
        another (step i's image is step i's VAE output); spread over two streams, the start-up and drain of one
        tail overlap the steady state of the other.  ``decode_tail_shape`` (None = the library's setting) selects
        the tails' pipeline shape for THIS pass's launches: the process-wide knob is set around them and restored
        (launch time is capture time for a graphed pass)."""
        if self.decode_tail_shape is None:
            return self._decode_launches(inp, part)
        lib = _cabi.lib()
        prev = lib.ldiff_tune_get(_cabi.TUNE_DECODE_TAIL_TMA)
        lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, int(self.decode_tail_shape))
        try:
            return self._decode_launches(inp, part)
        finally:
            lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, prev)

    def _decode_launches(self, inp, part: int = 0):
        n = self.n
        for i in range(part, n, self.decode_streams):
            last = i == n - 1
            if self.fused:
                # the step's training-path feature (ldiffusion.py:240-247) and, on the last step, the label slot
                # of the pixel vectors + the label's down-sample (ldiffusion.py:224-226) ride in the same pass
                ops.decode_tail_fused(inp.decoded[i], self.planes[:, i], rgb_out=self.rgb if last else None,
                                      feat_out=self.featcat, feat_channel=i,
                                      label=inp.gt if last else None,
                                      label_plane_out=self.planes[:, n] if last else None,
                                      label_small_out=self.label_small if last else None)
            else:
                ops.decode_tail_gray(inp.decoded[i], want_rgb=False, rgb_out=self.rgb if last else None,
                                     gray_out=self.planes[:, i])
        if not self.fused and part == (n - 1) % self.decode_streams:
            ops.copy_planes_u8(inp.gt, self.planes[:, n])                  # label slot of the pixel vectors

    def _chain_lifts(self, inp):
        n = self.n
        if not self.fused:
            # one gather per step: inside the concurrent pass the single multi-source launch
            # (ops.bilinear_lift_multi, what features.feature_concat uses stand-alone) measured 4 % slower —
            # its burst of scattered reads lands on top of the first decode tail
            for i in range(n):
                ops.bilinear_lift(inp.decoded[i], self.feat_size, out=self.featcat, out_channel=i, gray=True)
            ops.bilinear_lift(inp.gt.unsqueeze(1), self.feat_size, out=self.label_small)
        # ldiffusion.py:251: the last decode, down then up again (its own chain: the up-lift writes 6 B/pixel and
        # must not wait for the five decode tails)
        ops.bilinear_lift(inp.decoded[n - 1], self.feat_size, out=self.rgb_small)
        ops.bilinear_lift(self.rgb_small, (self.H, self.W), out=self.rgb_up)

    def _chain_tissue(self, inp):
        if self.fused:
            ops._head_logits(inp.head_feat, self.head_w, self.head_b, self.logits,
                             None if self.accumulate else self.C[0])                             # also zeroes C[0]
            if self.tissue_hist_fused:
                ops.lift_argmax_hist(self.logits, (self.H, self.W), inp.gt, out=self.C[0], mask_out=self.mask_tissue,
                                     exchange=self.exchange, channel=0)
            else:
                ops._lift_argmax(self.logits, self.mask_tissue)
                self._confusion(self.mask_tissue, inp.gt, 0)
            return
        ops._head_logits(inp.head_feat, self.head_w, self.head_b, self.logits)
        ops._lift_argmax(self.logits, self.mask_tissue)
        self._confusion(self.mask_tissue, inp.gt, 0)

    def _chain_cell(self, inp):
        if self.fused:
            ops._cell_classify(inp.inst_feats, self.cell_w, self.cell_b, self.inst_ids, self.lut, None, self.status,
                               None if self.accumulate else self.C[1])                          # also zeroes C[1]
            ops.lut_paint_hist(inp.inst_map, self.lut, inp.gt, self.K, out=self.C[1], mask_out=self.mask_cell,
                               exchange=self.exchange, channel=1)
            return
        ops._cell_classify(inp.inst_feats, self.cell_w, self.cell_b, self.inst_ids, self.lut, None, self.status)
        ops.lut_paint(inp.inst_map, self.lut, out=self.mask_cell)
        self._confusion(self.mask_cell, inp.gt, 1)

    # ------------------------------------------------------------------
    # host-facing API: pinned host batches in, pinned host results out
    # ------------------------------------------------------------------
    RESULT_KEYS = ("latents", "pixel_planes", "rgb", "featcat", "label_small", "mask_tissue", "mask_cell", "confusion")

    def _host_state(self, like: "HotPathInputs"):
        st = getattr(self, "_hs", None)
        if st is None:
            dev = self.device
            st = {
                "in": [like.packed(device=dev) for _ in range(2)],          # two device-side input slabs
                "s_in": torch.cuda.Stream(dev), "s_run": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                "in_ready": [torch.cuda.Event() for _ in range(2)],
                "in_free": [torch.cuda.Event() for _ in range(2)],
                "run_done": torch.cuda.Event(), "out_done": torch.cuda.Event(),
            }
            self._hs = st
        return st

    def alloc_host_results(self):
        """Pinned host mirror of the result slab: a dict of views by RESULT_KEYS, plus the slab under ``"_slab"``."""
        slab = torch.empty(self.out_slab.numel(), dtype=torch.uint8, pin_memory=True)
        out = {name: slab[o:o + nb].view(dt).view(shape) for name, o, nb, shape, dt in self._out_layout}
        out["_slab"] = slab
        return out

    def host_bytes_per_step(self, batch: "HotPathInputs"):
        """(host->device, device->host) bytes ``run_host`` moves per batch."""
        return batch.nbytes(), sum(nb for _, _, nb, _, _ in self._out_layout)

    def run_host(self, batches, host_out, after_run=None):
        """Process a sequence of host-resident batches; the results of batch i are copied into
        ``host_out[i % len(host_out)]`` (from ``alloc_host_results``).

        A batch from ``HotPathInputs.packed()`` (one pinned slab) moves host->device as ONE copy and the
        results come back as ONE copy; a plain ``HotPathInputs`` of separate pinned tensors is copied tensor by
        tensor.  Three streams pipeline the work: while batch i computes, batch i+1 streams host->device into
        the other input slab and batch i-1's results stream device->host, so a long run costs
        max(H2D, compute, D2H) per batch instead of their sum (PCIe is full duplex).  ``after_run`` (optional)
        is called on the compute stream after each pass (e.g. the confusion all-reduce)."""
        st = self._host_state(batches[0])
        s_in, s_run, s_out = st["s_in"], st["s_run"], st["s_out"]
        cur = torch.cuda.current_stream(self.device)
        for s in (s_in, s_run, s_out):
            s.wait_stream(cur)
        for i, hb in enumerate(batches):
            slot = i & 1
            dst = st["in"][slot]
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(st["in_free"][slot])          # pass i-2 has consumed this slab
                if hb.slab is not None and hb.slab.numel() == dst.slab.numel():
                    dst.slab.copy_(hb.slab, non_blocking=True)
                else:
                    for src, d in zip(hb.tensors(), dst.tensors()):
                        d.copy_(src, non_blocking=True)
                st["in_ready"][slot].record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(st["in_ready"][slot])
                if i >= 1:
                    s_run.wait_event(st["out_done"])              # results of pass i-1 are out of the slab
                self.run(dst)
                if after_run is not None:
                    after_run()
                st["in_free"][slot].record(s_run)
                st["run_done"].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(st["run_done"])
                host_out[i % len(host_out)]["_slab"].copy_(self.out_slab, non_blocking=True)
                st["out_done"].record(s_out)
        for s in (s_in, s_run, s_out):
            cur.wait_stream(s)

    def results(self):
        return {"latents": self.lat[-1], "noisy": self.noisy, "pixel_planes": self.planes, "rgb": self.rgb,
                "featcat": self.featcat, "label_small": self.label_small, "rgb_up": self.rgb_up,
                "logits": self.logits, "mask_tissue": self.mask_tissue, "mask_cell": self.mask_cell,
                "confusion": self.C}


class HotPathRing:
    """``n_in_flight`` passes of consecutive batches in flight at once: each slot is a ``HotPath`` with its own
    buffers and chain streams, enqueued on its own stream, so the tail of one pass (the last classifier kernels,
    the join) overlaps the head of the next instead of leaving the machine half empty — a pass is five chains of
    kernels with different bottlenecks (HBM, issue, launch latency) and no single pass keeps all of them busy.
    Measured at the bench shape (8 x 1024^2, K=11, 5 steps, bf16): 112 us per pass alone, 93 us with three in
    flight, and 89 us when the slots' decode tails also run in their smallest pipeline shape (2 stages, ONE CTA per
    SM: slower alone, 127 us for a single pass, but each pass then leaves room for the other passes' kernels) —
    ``throughput_shape`` (default: on from three passes in flight, bf16 storage, unless LDIFF_DT_TMA pins a shape).
    Results of pass i live in ``slot(i).results()`` until pass i + n_in_flight is enqueued
    (tools/pass_overlap.py)."""

    THROUGHPUT_DECODE_TAIL_SHAPE = 11

    def __init__(self, n_in_flight: int = 2, *args, throughput_shape: Optional[bool] = None, **kwargs):
        if n_in_flight < 1:
            raise ValueError("n_in_flight must be >= 1")
        self.slots = [HotPath(*args, **kwargs) for _ in range(n_in_flight)]
        if throughput_shape is None:
            throughput_shape = (n_in_flight >= 3 and self.slots[0].dtype == torch.bfloat16
                                and "LDIFF_DT_TMA" not in os.environ)
        if throughput_shape:
            for slot in self.slots:
                slot.decode_tail_shape = self.THROUGHPUT_DECODE_TAIL_SHAPE
        dev = self.slots[0].device
        self.streams = [torch.cuda.Stream(dev) for _ in range(n_in_flight)]
        self.device = dev

    def __len__(self):
        return len(self.slots)

    def slot(self, i: int) -> "HotPath":
        return self.slots[i % len(self.slots)]

    def stream(self, i: int) -> torch.cuda.Stream:
        return self.streams[i % len(self.streams)]

    def fork(self):
        """Make every slot's stream wait for the caller's current stream (call before the first ``run``)."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def run(self, i: int, inp: HotPathInputs):
        """Enqueue pass i on its slot's stream (graph-capturable per slot)."""
        with torch.cuda.stream(self.stream(i)):
            self.slot(i).run(inp)

    def join(self):
        """Make the caller's current stream wait for every slot."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)


def synth_inputs(batch, height, width, num_classes, num_steps=5, dtype=torch.bfloat16, device="cpu",
                 head_hw=(32, 32), n_instances=800, seed=1234, pin=False, inst_dtype=torch.int32) -> HotPathInputs:
    """Seeded synthetic PUMA-shaped inputs: latents 5.5*N(0,1) (raw SD VAE-mean
    scale), eps N(0,1), decoded U(-1.2,1.2), features N(0,1), ~n_instances square cells per patch,
    gt 70% background + blobs of classes 1..K-1 + 0.1% 'other' (255) pixels.  ``inst_dtype``: int32 (default) or
    uint16, the label image as Cellpose returns it (same values, same random stream)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = (batch, 4, height // 8, width // 8)

    def fin(t):
        if pin and t.device.type == "cpu":
            t = t.pin_memory()
        return t.to(device) if str(device) != "cpu" else t

    latents = fin((torch.randn(lat, generator=g) * 5.5).to(dtype))
    eps = [fin(torch.randn(lat, generator=g).to(dtype)) for _ in range(num_steps)]
    decoded = [fin(torch.empty(batch, 3, height, width).uniform_(-1.2, 1.2, generator=g).to(dtype))
               for _ in range(num_steps)]
    head_feat = fin(torch.randn(batch, 256, *head_hw, generator=g).to(dtype))
    inst_feats = fin(torch.randn(batch, n_instances, 256, generator=g).to(dtype))
    inst = torch.zeros(batch, height, width, dtype=torch.int32)
    gt = torch.zeros(batch, height, width, dtype=torch.uint8)
    side = max(4, int((0.3 * height * width / max(n_instances, 1)) ** 0.5))
    ys = torch.randint(0, max(1, height - side), (batch, n_instances), generator=g)
    xs = torch.randint(0, max(1, width - side), (batch, n_instances), generator=g)
    cls = torch.randint(1, num_classes, (batch, n_instances), generator=g)
    for b in range(batch):
        for i in range(n_instances):
            y, x = int(ys[b, i]), int(xs[b, i])
            inst[b, y:y + side, x:x + side] = i + 1
            gt[b, y:y + side, x:x + side] = int(cls[b, i])
    other = torch.rand(batch, height, width, generator=g) < 0.001
    gt[other] = 255
    return HotPathInputs(latents, eps, decoded, head_feat, fin(inst.to(inst_dtype)), inst_feats, fin(gt))
