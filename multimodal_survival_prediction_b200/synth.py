"""Seeded synthetic cohorts shaped like the reference's data.

``time`` are integer day counts like days_to_death / days_to_last_follow_up
(scripts/preprocessing/create_full_matching_table.py:52-55): clamp(floor(Exp(mean 1000)), 1, 4000),
about 30 % events, log-hazards N(0,1).  Generated on the CPU generator so that every rank and the
oracle see identical values for a given seed.
"""
from __future__ import annotations

import torch


def cohort(n: int, seed: int = 0, event_rate: float = 0.30, few_ties: bool = False,
           risk_tie_frac: float = 0.0, chunk: int = 1 << 22):
    """Return (log_hz f32[n], event bool[n], time f32[n]) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    log_hz = torch.empty(n, dtype=torch.float32)
    time = torch.empty(n, dtype=torch.float32)
    event = torch.empty(n, dtype=torch.bool)
    for a in range(0, n, chunk):  # chunked so that 16M rows do not need float64 temporaries at once
        b = min(n, a + chunk)
        t = torch.empty(b - a, dtype=torch.float32).exponential_(1.0 / 1000.0, generator=g)
        if not few_ties:
            t = torch.clamp(torch.floor(t), 1.0, 4000.0)
        time[a:b] = t
        event[a:b] = torch.rand(b - a, generator=g) < event_rate
        log_hz[a:b] = torch.randn(b - a, generator=g)
    if risk_tie_frac > 0:
        sel = torch.rand(n, generator=g) < risk_tie_frac
        log_hz[sel] = torch.round(log_hz[sel] * 100) / 100
    return log_hz, event, time


def modality_batch(batch: int, rna_dim: int = 5005, seed: int = 0, ct_feat_dim: int = 128):
    """Head inputs with the 608-cohort's modality availability rates
    (results/final_comparison/results.json: imaging 142, RNA-seq 427, clinical 587 of 608);
    missing modalities are zero-filled as the dataset does (partial_modality_training.py:89,116,125)."""
    g = torch.Generator().manual_seed(seed)
    mask = torch.stack([torch.rand(batch, generator=g) < 142 / 608,
                        torch.rand(batch, generator=g) < 427 / 608,
                        torch.rand(batch, generator=g) < 587 / 608], dim=1).float()
    ct_feat = torch.relu(torch.randn(batch, ct_feat_dim, generator=g)) * mask[:, 0:1]
    rna = torch.randn(batch, rna_dim, generator=g) * mask[:, 1:2]
    clinical = (0.3 + 0.6 * torch.rand(batch, 1, generator=g)) * mask[:, 2:3]
    return ct_feat, rna, clinical, mask
