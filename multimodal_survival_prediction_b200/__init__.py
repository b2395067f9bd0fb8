"""B200-native survival-training hot path: Cox partial likelihood, Harrell's C-index and the gated
late-fusion head of baek0203/multimodal_survival_prediction, as hand-written sm_100a CUDA behind the
C ABI in include/b200surv.h.  Importing the package does not need a GPU; calling an operator does
(there is no CPU fallback)."""
from . import _lib
from ._lib import B200SurvError
from .cindex import ConcordanceIndex, cindex_counts, cindex_from_counts
from .compact import ValidationCohort, select_labelled
from .cox import neg_partial_log_likelihood, neg_partial_log_likelihood_segmented

__all__ = ["B200SurvError", "ConcordanceIndex", "cindex_counts", "cindex_from_counts",
           "neg_partial_log_likelihood", "neg_partial_log_likelihood_segmented", "select_labelled", "ValidationCohort"]
