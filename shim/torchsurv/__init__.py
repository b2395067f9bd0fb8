"""Import shim: lets the reference's training scripts run unchanged against the B200 kernels.

The scripts do ``from torchsurv.loss.cox import neg_partial_log_likelihood`` and
``from torchsurv.metrics.cindex import ConcordanceIndex`` inside try/except ImportError
(scripts/training/simple_fusion.py:22-29).  Put this directory ahead of site-packages
(``PYTHONPATH=<repo>/shim:<repo>``) and those imports resolve to
multimodal_survival_prediction_b200.  This is NOT torchsurv: only the two entry points the
reference uses exist."""
import os as _os

if _os.environ.get("B200SURV_NO_TORCHSURV_SHIM"):
    # lets a harness drive the reference scripts down their OTHER branch (in-repo loss + lifelines C-index,
    # partial_modality_training.py:295-319) with only the lifelines shim active
    raise ImportError("torchsurv shim disabled by B200SURV_NO_TORCHSURV_SHIM")

__version__ = "0.0+b200surv"
