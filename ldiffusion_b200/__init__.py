"""ldiffusion_b200 — B200-native (sm_100a) implementation of the L-Diffusion
sampling-and-feature hot path behind the reference's Python call signatures.

Public surface (drop-in seams, see DESIGN.md / INTEGRATION.md):

* ``LaplacePLMSScheduler``            - ``pipeline.scheduler`` duck-type + Laplace q_sample
* ``pixel_latent_vector`` / ``PixelVectorBuilder`` / ``feature_concat``
* ``TissueHead`` / ``cell_mask``      - classifier head + argmax
* ``micro_dice`` / ``mean_iou_and_per_class`` / ``pixel_accuracy`` /
  ``frequency_weighted_iou`` / ``evaluate``
* ``pixel_contrastive_loss``          - InfoNCE term of the warm-up loss (forward + backward)
* ``Segmentor`` / ``LDiffusionModel`` - orchestrator shims with the reference's signatures
* ``dataset`` / ``utils``             - label tables as LUTs, dataset-preparation callers of the loop

All tensor work runs in ``libldiff_sm100.so`` (hand-written CUDA behind the C ABI
of ``include/ldiff.h``); there is no CPU fallback.
"""
from . import ops  # noqa: F401  (registers the ldiff:: custom ops)
from . import dataset, utils  # noqa: F401
from .features import (PixelVectorBuilder, feature_concat, label_down, pixel_latent_vector,  # noqa: F401
                       pixel_vectors, rgb_up)
from .head import TissueHead, cell_mask, tissue_mask  # noqa: F401
from .metrics import (confusion_matrix, evaluate, frequency_weighted_iou, mean_iou_and_per_class,  # noqa: F401
                      micro_dice, pixel_accuracy)
from .ldiffusion import LDiffusionModel  # noqa: F401
from .loss import pixel_contrastive_loss, sample_contrastive_pairs, sample_contrastive_pairs_device  # noqa: F401
from .scheduler import LaplacePLMSScheduler  # noqa: F401
from .segmentor import Segmentor  # noqa: F401

__version__ = "0.1.0"
