"""On-disk artefacts interchangeable with the reference.

* checkpoints: the reference saves ``model.state_dict()`` per fold when the validation C-index improves
  (scripts/training/partial_modality_training.py:553-556 -> ``models/partial_modality/fold_{k}_best.pth``;
  simple_fusion.py:404-406).  ``head.PartialModalityNet`` / ``MultiModalSurvivalNet`` keep the reference's keys and
  shapes, so ``save_fold_checkpoint`` / ``load_fold_checkpoint`` are torch.save / load_state_dict(strict=True) on CPU
  tensors, readable by either side.
* ``cv_results.json``: the summary the K-fold drivers write (partial_modality_training.py:592-607) and the analysis layer
  reads (scripts/training/final_comparison.py:42-58, scripts/analysis/analyze_all_results.py:34-64): ``c_index_mean``,
  ``c_index_std`` (numpy population std) and ``fold_results[*].best_c_index``.
"""
from __future__ import annotations

import json
import math
import os

import torch


def save_fold_checkpoint(model: torch.nn.Module, directory: str, fold: int) -> str:
    """``torch.save(model.state_dict(), f'{directory}/fold_{fold}_best.pth')`` with CPU tensors (fold is 1-based)."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"fold_{fold}_best.pth")
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
    return path


def load_fold_checkpoint(model: torch.nn.Module, path: str):
    """strict=True: a checkpoint written by the reference's class must match key for key."""
    sd = torch.load(path, map_location="cpu")
    return model.load_state_dict(sd, strict=True)


def write_cv_results(path: str, model: str, fold_results, hyperparameters=None, **extra) -> dict:
    """Write the reference's summary schema.  ``fold_results``: list of dicts with at least ``fold`` and
    ``best_c_index`` (the drivers add ``train_size``, ``train_survival_size``, ``val_size`` / ``best_epoch``)."""
    vals = [float(r["best_c_index"]) for r in fold_results]
    if not vals or not all(math.isfinite(v) for v in vals):
        raise ValueError("fold_results need finite best_c_index values")
    mean = sum(vals) / len(vals)
    std = math.sqrt(sum((v - mean) ** 2 for v in vals) / len(vals))      # np.std: population standard deviation
    summary = {"model": model, "c_index_mean": mean, "c_index_std": std,
               "fold_results": [dict(r, best_c_index=float(r["best_c_index"])) for r in fold_results]}
    if hyperparameters is not None:
        summary["hyperparameters"] = dict(hyperparameters)
    summary.update(extra)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(summary, fh, indent=2)
    return summary


def read_cv_results(path: str) -> dict:
    """The analysis layer's reading rule (final_comparison.py:42-58): mean/std from the file when present, else from the
    folds; ``fold_values`` always from ``fold_results[*].best_c_index``."""
    with open(path) as fh:
        data = json.load(fh)
    folds = [f["best_c_index"] for f in data["fold_results"]]
    if "c_index_mean" in data:
        mean, std = data["c_index_mean"], data.get("c_index_std", 0)
    else:
        mean = sum(folds) / len(folds)
        std = math.sqrt(sum((v - mean) ** 2 for v in folds) / len(folds))
    return {"mean": mean, "std": std, "fold_values": folds}
