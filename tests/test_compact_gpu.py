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


def test_validation_cohort_matches_reference_validate_loop():
    """ValidationCohort against the accumulation of the reference's validate(), partial_modality_training.py:438-485:
    boolean indexing per batch, the skip rule n >= 2 and events > 0, host lists, C-index over everything kept."""
    from multimodal_survival_prediction_b200 import ConcordanceIndex, ValidationCohort, neg_partial_log_likelihood
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7)
    sizes = [4, 4, 1, 4, 3, 64, 2, 4, 517, 4]       # the reference validates with batch 4; ragged tail batches
    batches = []
    for i, B in enumerate(sizes):
        hazard = torch.randn(B, generator=g)
        time = torch.floor(torch.rand(B, generator=g) * 400) + 1
        event = (torch.rand(B, generator=g) < (0.0 if i == 3 else 0.4)).float()     # batch 3: no event -> skipped
        has = (torch.rand(B, generator=g) < (0.0 if i == 1 else 0.6)).tolist()      # batch 1: nothing labelled
        batches.append((hazard.to(dev), torch.stack([time, event], 1).to(dev), has))

    def cox_loss(h, e, t):
        return neg_partial_log_likelihood(h, e.bool(), t)

    cohort = ValidationCohort(capacity=sum(sizes), device=dev)
    total, nb, hs, ts, es = 0.0, 0, [], [], []
    for hazard, label, has in batches:
        added = cohort.add(hazard, label, has, loss_fn=cox_loss)
        m = torch.tensor(has, dtype=torch.bool, device=dev)
        k = 0
        if m.sum() > 0:
            h, t, e = hazard[m], label[m, 0], label[m, 1]
            if h.shape[0] >= 2 and e.sum() > 0:
                total += cox_loss(h, e, t).item(); nb += 1; k = h.shape[0]
                hs.extend(h.cpu().numpy()); ts.extend(t.cpu().numpy()); es.extend(e.cpu().numpy())
        assert added == k
    assert nb >= 5 and cohort.num_batches == nb and cohort.n == len(hs)
    h, e, t = cohort.vectors()
    assert torch.equal(h.cpu(), torch.tensor(hs)) and torch.equal(t.cpu(), torch.tensor(ts))
    assert torch.equal(e.cpu(), torch.tensor(es).bool())
    avg, c = cohort.finish()
    ref_c = ConcordanceIndex()(torch.tensor(hs), torch.tensor(es).bool(), torch.tensor(ts)).item()
    assert abs(avg - total / nb) <= 1e-6 * abs(total / nb) and c == ref_c
    empty = ValidationCohort(capacity=8, device=dev)
    assert empty.finish() == (0, 0.5)
    with pytest.raises(ValueError):
        empty.add(torch.zeros(9, device=dev), torch.zeros(9, 2, device=dev), [True] * 9)
