"""TEST/BENCH INFRASTRUCTURE -- PyTorch-CPU restatement of the reference's ungated late-fusion net and of one training
step of it (BASELINE.json configs[0]: simple_fusion.py / final_multimodal.py, batch 4, Cox loss on CPU).

Follows scripts/training/final_multimodal.py:59-150 (model), :158-162 (loss call) and simple_fusion.py:255-275
(step: forward, Cox loss, backward, clip_grad_norm_(1.0), AdamW step).  Only tests/ and bench.py's CPU baseline leg may
import it; the product path never does.  torchsurv is absent, so the loss is oracle/cox_torch.py (the vectorised CPU
port) wrapped in an autograd Function.
"""
import torch
from torch import nn

from . import cox_torch
from .ctenc import reference_cnn


class MultiModalNetCPU(nn.Module):
    def __init__(self, rna_dim=5005, clinical_dim=1):
        super().__init__()
        self.ct_encoder = reference_cnn()
        self.rna_encoder = nn.Sequential(nn.Linear(rna_dim, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.3),
                                         nn.Linear(512, 128), nn.ReLU())
        self.clinical_encoder = nn.Sequential(nn.Linear(clinical_dim, 32), nn.ReLU())
        self.fusion = nn.Sequential(nn.Linear(128 + 128 + 32, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.3),
                                    nn.Linear(256, 128), nn.ReLU())
        self.cox_head = nn.Linear(128, 1)

    def forward(self, ct, rna, clinical):
        feats = [self.ct_encoder(ct).view(ct.size(0), -1), self.rna_encoder(rna), self.clinical_encoder(clinical)]
        return self.cox_head(self.fusion(torch.cat(feats, dim=1))).squeeze(1)


class _CoxCPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_hz, event, time):
        loss, grad = cox_torch.cox_nll_fwd_bwd(log_hz.detach(), event, time)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        return g * ctx.saved_tensors[0], None, None


def training_step(model, optimizer, ct, rna, clinical, event, time):
    """One step of simple_fusion.py:255-275 on CPU tensors; returns the loss value."""
    optimizer.zero_grad()
    loss = _CoxCPU.apply(model(ct, rna, clinical), event, time)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    optimizer.step()
    return float(loss.detach())
