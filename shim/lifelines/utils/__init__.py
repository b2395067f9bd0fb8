"""lifelines.utils shim -> B200 kernels (see shim/lifelines/__init__.py)."""
from multimodal_survival_prediction_b200.cindex import concordance_index_lifelines as concordance_index  # noqa: F401
