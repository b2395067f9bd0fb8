"""torchsurv.metrics.cindex shim -> B200 kernels (see shim/torchsurv/__init__.py)."""
from multimodal_survival_prediction_b200.cindex import ConcordanceIndex  # noqa: F401
