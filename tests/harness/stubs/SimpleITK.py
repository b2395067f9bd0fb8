"""Stub: the reference's training scripts do an unconditional ``import SimpleITK as sitk`` (simple_fusion.py:42,
partial_modality_training.py:57) but only call it for rows whose ``nifti_path`` exists; the harness cohort has none."""
