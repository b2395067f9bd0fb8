"""torchsurv.loss.cox shim -> B200 kernels (see shim/torchsurv/__init__.py)."""
from multimodal_survival_prediction_b200.cox import neg_partial_log_likelihood  # noqa: F401
