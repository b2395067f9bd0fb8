"""Import shim: ``from lifelines.utils import concordance_index`` (the reference's second C-index fallback,
scripts/training/partial_modality_training.py:313-319, scripts/analysis/evaluate_model.py:24,41-45) resolves to
the B200 pair-count kernel.  This is NOT lifelines: only that one function exists."""
__version__ = "0.0+b200surv"
