"""Fused clip_grad_norm_ + Adam / AdamW against torch's own pair, the one the reference calls at
scripts/training/partial_modality_training.py:427-428 (Adam, weight_decay 1e-4) and simple_fusion.py:273-274 (AdamW)."""
import pytest
import torch

from multimodal_survival_prediction_b200.optim import ClipAdam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("adamw", [False, True])
@pytest.mark.parametrize("max_norm", [1.0, 0.0])
def test_clip_adam_matches_torch(adamw, max_norm):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    shapes = [(512, 5005), (512,), (128, 512), (3, 64), (1,), (4097,), (32, 1)]
    ours = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4 if not adamw else 1e-2)
    opt = ClipAdam(ours, max_norm=max_norm, adamw=adamw, **kw)
    topt = (torch.optim.AdamW if adamw else torch.optim.Adam)(ref, **kw)
    for it in range(5):
        for p, q in zip(ours, ref):
            g = torch.randn_like(p) * (3.0 if it % 2 else 0.01)      # clipped and unclipped steps
            p.grad = g.clone(); q.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm > 0 else None
        topt.step()
        opt.step()
        if tn is not None:
            assert abs(float(opt.last_total_norm) - float(tn)) <= 1e-5 * float(tn)
        for p, q in zip(ours, ref):
            assert float((p.detach() - q.detach()).abs().max()) <= 2e-6 * max(1.0, float(q.detach().abs().max())), (it, p.shape)
