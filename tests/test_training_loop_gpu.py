"""End-to-end on a B200: the reference's training step (scripts/training/partial_modality_training.py:382-435 --
model(ct_feat, rna, clinical, mask) -> labelled-row selection -> cox_loss -> + 0.01 * gate entropy -> backward ->
clip_grad_norm_ -> Adam) and its validation pass (:438-485, ConcordanceIndex on CPU tensors), with the three operators
coming from this package through the torchsurv shim's import path.  The synthetic cohort carries signal in the RNA
block, so the loss must fall and the C-index must rise well above chance."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused_step", [False, True])
def test_reference_style_training_loop_learns(fused_step):
    """fused_step: the labelled-row selection and clip + Adam also come from this package (select_labelled, ClipAdam)."""
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    try:
        from torchsurv.loss.cox import neg_partial_log_likelihood        # resolves to the B200 operators
        from torchsurv.metrics.cindex import ConcordanceIndex
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))
    from multimodal_survival_prediction_b200 import head as ghead
    from multimodal_survival_prediction_b200 import synth

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    n, B = 2048, 256
    ct, rna, clin, mask = synth.modality_batch(n, seed=11)
    beta = torch.zeros(rna.shape[1]); beta[:20] = 0.5
    risk = (rna * mask[:, 1:2]) @ beta                                   # only rows that have RNA-seq carry signal
    g = torch.Generator().manual_seed(5)
    t_event = torch.empty(n).exponential_(1.0, generator=g) * torch.exp(-risk) * 1000.0
    t_cens = torch.empty(n).exponential_(1.0 / 1500.0, generator=g)
    time = torch.clamp(torch.floor(torch.minimum(t_event, t_cens)), 1, 8000)
    event = (t_event <= t_cens)
    has_survival = torch.rand(n, generator=g) < 348 / 608
    label = torch.stack([time, event.float()], 1)

    model = ghead.PartialModalityNet().to(dev)
    from multimodal_survival_prediction_b200.compact import select_labelled
    from multimodal_survival_prediction_b200.optim import ClipAdam
    params = [p for n_, p in model.named_parameters() if not n_.startswith("ct_encoder")]
    opt = (ClipAdam(params, lr=1e-3, weight_decay=1e-4, max_norm=1.0) if fused_step
           else torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4))

    def cox_loss(hazard, ev, tm):                                         # partial_modality_training.py:285-288
        return neg_partial_log_likelihood(hazard, ev.bool(), tm)

    def epoch(train):
        model.train(train)
        tot, nb = 0.0, 0
        hz_all, ev_all, t_all = [], [], []
        for a in range(0, n, B):
            sl = slice(a, a + B)
            c, r, cl, m = (x[sl].to(dev) for x in (ct, rna, clin, mask))
            lab, surv = label[sl].to(dev), has_survival[sl].to(dev)
            with torch.set_grad_enabled(train):
                hazard, gate = model.forward_features(c, r, cl, m)
                if fused_step:
                    hs, ts, es, n_ev = select_labelled(hazard, lab, surv)
                else:
                    hs, ts, es = hazard[surv], lab[surv, 0], lab[surv, 1]
                    n_ev = int(es.sum())
                c_loss = cox_loss(hs, es, ts) if hs.shape[0] >= 2 and n_ev > 0 else torch.tensor(0.0, device=dev)
                loss = c_loss + 0.01 * ghead.gate_entropy_loss(gate)
            if train:
                opt.zero_grad()
                loss.backward()
                if not fused_step:
                    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()
            tot += float(c_loss.detach()); nb += 1
            hz_all.append(hs.detach().cpu()); ev_all.append(es.detach().cpu()); t_all.append(ts.detach().cpu())
        ci = ConcordanceIndex()(torch.cat(hz_all), torch.cat(ev_all).bool(), torch.cat(t_all))   # CPU tensors, like :478-481
        return tot / nb, float(ci)

    loss0, ci0 = epoch(False)
    for _ in range(6):
        tr_loss, _ = epoch(True)
    loss1, ci1 = epoch(False)
    assert torch.isfinite(torch.tensor([loss0, loss1, tr_loss])).all()
    assert loss1 < loss0 - 0.05, (loss0, loss1)
    assert ci1 > 0.62 and ci1 > ci0 + 0.05, (ci0, ci1)
