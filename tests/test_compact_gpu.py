"""Labelled-row compaction against the reference's boolean indexing,
scripts/training/partial_modality_training.py:401-408."""
import pytest
import torch

from multimodal_survival_prediction_b200.compact import select_labelled

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 4, 37, 2048, 2049, 4096, 100_003])
@pytest.mark.parametrize("p_keep", [0.0, 348 / 608, 1.0])
def test_select_labelled_matches_boolean_indexing(B, p_keep):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(B)
    hazard = torch.randn(B, generator=g).to(dev).requires_grad_(True)
    h2 = hazard.detach().clone().requires_grad_(True)
    label = torch.stack([torch.rand(B, generator=g) * 4000, (torch.rand(B, generator=g) < 0.3).float()], 1).to(dev)
    has_survival = (torch.rand(B, generator=g) < p_keep).tolist()       # a python list, like the reference's batch field
    hs, ts, es, n_ev = select_labelled(hazard, label, has_survival)
    mask = torch.tensor(has_survival, dtype=torch.bool, device=dev)    # partial_modality_training.py:401-406
    rh, rt, re = h2[mask], label[mask, 0], label[mask, 1]
    assert torch.equal(hs, rh) and torch.equal(ts, rt) and torch.equal(es, re.bool())
    assert n_ev == int(re.sum())
    w = torch.randn(hs.shape[0], device=dev)
    (hs * w).sum().backward()
    (rh * w).sum().backward()
    assert torch.equal(hazard.grad, h2.grad)
