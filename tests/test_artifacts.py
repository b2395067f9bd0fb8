"""``.pth`` and ``cv_results.json`` interchange with the reference (CPU: no kernels involved).

The reference's class is AST-extracted from its script (tests/_ref_scripts/, placed by tests/harness/prepare_ref_scripts.py)
and its ``state_dict`` goes through torch.save / torch.load into the B200 head with strict=True, and back."""
import ast
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import prepare_ref_scripts  # noqa: E402

from multimodal_survival_prediction_b200 import artifacts  # noqa: E402
from multimodal_survival_prediction_b200 import head as ghead  # noqa: E402


def reference_class():
    prepare_ref_scripts.prepare()
    script = os.path.join(ROOT, "tests", "_ref_scripts", "partial_modality_training.py")
    if not os.path.exists(script):
        pytest.skip("tests/_ref_scripts/ is empty (no /root/reference here)")
    tree = ast.parse(open(script, encoding="utf-8").read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PartialModalityNet"][0]
    ns = {"torch": torch, "nn": torch.nn, "USE_MONAI": False}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), script, "exec"), ns)
    return ns["PartialModalityNet"]


def test_pth_round_trip_strict_both_ways(tmp_path):
    Ref = reference_class()
    torch.manual_seed(3)
    ref = Ref()
    ref.rna_encoder[1].running_mean.normal_()           # non-default buffers must travel too
    ref.fusion[1].num_batches_tracked.fill_(17)
    os.makedirs(tmp_path / "models" / "partial_modality")
    p = tmp_path / "models" / "partial_modality" / "fold_1_best.pth"
    torch.save(ref.state_dict(), p)                     # partial_modality_training.py:555
    ours = ghead.PartialModalityNet()
    res = artifacts.load_fold_checkpoint(ours, str(p))
    assert res.missing_keys == [] and res.unexpected_keys == []
    for k, v in ref.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v), k
    p2 = artifacts.save_fold_checkpoint(ours, str(tmp_path / "models" / "b200"), 2)
    assert p2.endswith("fold_2_best.pth")
    back = Ref()
    res = back.load_state_dict(torch.load(p2, map_location="cpu"), strict=True)
    assert res.missing_keys == [] and res.unexpected_keys == []
    assert list(back.state_dict()) == list(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(back.state_dict()[k], v) and back.state_dict()[k].dtype == v.dtype, k


def test_cv_results_json_schema(tmp_path):
    folds = [{"fold": i + 1, "best_c_index": c, "train_size": 538, "train_survival_size": 278, "val_size": 70}
             for i, c in enumerate([0.6081193089485168, 0.6057971119880676, 0.5627849102020264])]
    hp = {"batch_size": 8, "learning_rate": 1e-4, "epochs": 50, "n_folds": 3, "gate_entropy_weight": 0.01}
    path = tmp_path / "results" / "partial_modality" / "cv_results.json"
    artifacts.write_cv_results(str(path), "PartialModalityNet (Gating + Entropy Regularization)", folds, hp)
    data = json.load(open(path))
    # the key set and order of the reference's shipped results/partial_modality/cv_results.json
    assert list(data) == ["model", "c_index_mean", "c_index_std", "fold_results", "hyperparameters"]
    assert list(data["fold_results"][0]) == ["fold", "best_c_index", "train_size", "train_survival_size", "val_size"]
    vals = [f["best_c_index"] for f in folds]
    assert data["c_index_mean"] == pytest.approx(np.mean(vals), abs=1e-15) and data["c_index_std"] == pytest.approx(np.std(vals), abs=1e-15)
    got = artifacts.read_cv_results(str(path))           # the analysis layer's reading rule
    assert got["fold_values"] == vals and got["mean"] == data["c_index_mean"]
    with pytest.raises(ValueError):
        artifacts.write_cv_results(str(path), "x", [{"fold": 1, "best_c_index": float("nan")}])
