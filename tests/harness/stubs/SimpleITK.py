"""Stub: the reference's training scripts do an unconditional ``import SimpleITK as sitk`` (simple_fusion.py:42,
partial_modality_training.py:57) but only call it for rows whose ``nifti_path`` exists; the harness cohort has none.

The scripts never seed torch (weight init, DataLoader shuffles and dropout draw from the global generators), so an unchanged
run is not reproducible and its cross-validated C-index on the 60-patient harness cohort occasionally lands on the wrong side
of the assertions.  The harness seeds the generators here -- the one module the scripts import that the harness owns --
when B200SURV_HARNESS_SEED is set; the scripts themselves stay byte for byte the reference's."""
import os as _os

_seed = _os.environ.get("B200SURV_HARNESS_SEED")
if _seed is not None:
    import random as _random

    import numpy as _np
    import torch as _torch

    _random.seed(int(_seed))
    _np.random.seed(int(_seed))
    _torch.manual_seed(int(_seed))
