"""CPU oracle for the L-Diffusion sampling-and-feature hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  ``ldiffusion_b200`` never imports this package and has
no CPU fallback.

What it is: a restatement (torch-CPU / numpy / PIL, fp32) of the op chains the
reference executes on this path.  Every function cites the reference
``file:line`` it follows (paths relative to the reference checkout).

Two tiers live side by side:

* ``*_chain`` functions run the literal eager op chain of the reference
  (``F.interpolate``, ``torch.softmax``, ``torch.distributions.Laplace`` ...).
  They are what ``bench.py`` times as the CPU baseline.
* ``*_spec`` functions pin the *order of fp32 roundings* (no FMA contraction)
  so that the CUDA kernels can be compared bit for bit.  ``tests/`` checks each
  spec against its chain (exact where the chain is order-independent, within a
  stated ulp bound or "except near-ties" otherwise).

Pinning status
--------------
* metrics (a-6): PINNED against the reference's own ``utils.py`` /
  ``evaluate.py`` functions, imported in the build container through
  ``oracle/_refshim.py``; outputs committed as ``tests/golden/metrics_*.npz``
  by ``tests/golden/make_golden.py``.
* decode tail / PIL gray (a-3), bilinear (a-4), head + argmax (a-5), Laplace
  transform (a-1): the chain tier *is* the third-party code the reference
  calls (torch, numpy, PIL), so the spec tier is pinned against it here and
  through ``tests/golden/*.npz``.
* PNDM/PLMS scheduler (a-2): **parity unpinned**.  The algorithm lives in
  ``diffusers==0.34.0`` (pinned at ``environment.yml:42``), which is neither
  vendored in the reference nor installed here.  The restatement follows the
  published ``PNDMScheduler`` algorithm with the SD-v1.5 scheduler config and
  is anchored on the reference's call sites and on the known-answer constants
  recorded in ``SURVEY.md`` §8(a-2).
"""

F32 = "float32"
