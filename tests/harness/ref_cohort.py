"""Synthetic ``data/processed`` directory for the reference's training scripts (SURVEY.md 8b "Script harness").

``full_matching_table.csv`` has the columns written by scripts/preprocessing/create_full_matching_table.py:124-134;
``rnaseq_normalized_mapped.csv`` is indexed by patient_id with 5,005 z-scored gene columns (preprocess_genomic.py:108-117).
``nifti_path`` is empty for every row, so the datasets never call SimpleITK (partial_modality_training.py:91-92) and feed
zero volumes, as they do for patients without imaging.  Hazards carry signal: survival time shortens with the first genes.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd


def write(root: str, n_complete: int = 30, n_rna_only: int = 6, n_unlabelled: int = 12, rna_dim: int = 5005, seed: int = 0):
    rng = np.random.default_rng(seed)
    n = n_complete + n_rna_only + n_unlabelled
    ids = [f"TCGA-SY-{i:04d}" for i in range(n)]
    rna = rng.normal(size=(n, rna_dim)).astype(np.float32)
    risk = rna[:, :8].sum(axis=1) / np.sqrt(8.0)
    days = np.clip(np.floor(rng.exponential(900.0 * np.exp(-0.8 * risk))), 1, 4000)
    status = (rng.random(n) < 0.6).astype(int)
    has_survival = np.arange(n) < n_complete + n_rna_only
    has_imaging = np.arange(n) < n_complete
    has_rna = np.ones(n, bool)
    has_rna[n_complete + n_rna_only::3] = False                  # some unlabelled patients lack RNA-seq as well
    age = np.where(rng.random(n) < 0.95, rng.integers(30, 90, n).astype(float), np.nan)
    table = pd.DataFrame({
        "patient_id": ids, "nifti_path": [np.nan] * n, "has_imaging": has_imaging, "has_rnaseq": has_rna,
        "has_clinical": ~np.isnan(age), "age": age,
        "survival_time": np.where(has_survival, days, np.nan), "survival_status": np.where(has_survival, status, np.nan),
        "has_survival": has_survival})
    d = os.path.join(root, "data", "processed")
    os.makedirs(d, exist_ok=True)
    table.to_csv(os.path.join(d, "full_matching_table.csv"), index=False)
    genes = pd.DataFrame(rna[has_rna], index=pd.Index([i for i, h in zip(ids, has_rna) if h], name="patient_id"),
                         columns=[f"ENSG{j:011d}" for j in range(rna_dim)])
    genes.to_csv(os.path.join(d, "rnaseq_normalized_mapped.csv"))
    return table
