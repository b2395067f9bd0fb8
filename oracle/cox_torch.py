"""Multi-threaded PyTorch-CPU port of the Cox loss + gradient.  TEST/BENCH INFRASTRUCTURE ONLY.

This is the `--impl reference` arm of bench.py and the ``cpu_baseline`` leg: what the reference's
loss costs on the GPU box's host cores when written the way the reference (and torchsurv) write
it -- torch CPU ops in the input dtype (fp32) -- but vectorised, so that the baseline is not a Python
loop over distinct times.  torchsurv itself is absent from the image (parity unpinned,
oracle/__init__.py); the arithmetic follows oracle/cox.py, which is the float64 checker.
Reference call site: scripts/training/partial_modality_training.py:285-288.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def cox_nll_fwd_bwd(log_hz: torch.Tensor, event: torch.Tensor, time: torch.Tensor, ties_method: str = "efron",
                    reduction: str = "mean", efron_mean_over: str = "event_times"):
    """Returns (loss 0-dim, grad like log_hz).  CPU tensors; arithmetic in log_hz.dtype."""
    dt = log_hz.dtype
    n = log_hz.numel()
    ts, order = torch.sort(time, stable=True)
    es, ds = log_hz[order], event[order]
    _, gid, counts = torch.unique_consecutive(ts, return_inverse=True, return_counts=True)
    J = counts.numel()
    c = es.max()
    w = torch.exp(es - c)
    dsf = ds.to(dt)
    S = torch.zeros(J, dtype=dt).index_add_(0, gid, w)
    E = torch.zeros(J, dtype=dt).index_add_(0, gid, w * dsf)
    m = torch.zeros(J, dtype=torch.int64).index_add_(0, gid, ds.to(torch.int64))
    D = torch.flip(torch.cumsum(torch.flip(S, [0]), 0), [0])
    n_events = int(m.sum())
    if n_events == 0:
        return torch.zeros((), dtype=dt), torch.zeros_like(log_hz)
    has = m > 0
    if ties_method == "breslow":
        mm = m.to(dt)
        Dm = torch.where(has, D, torch.ones_like(D))
        T = torch.where(has, mm * (torch.log(Dm) + c), torch.zeros_like(D))
        G = torch.where(has, mm / Dm, torch.zeros_like(D))
        F = torch.zeros_like(D)
    else:
        g_rep = torch.repeat_interleave(torch.arange(J), m)
        start = torch.cumsum(m, 0) - m
        l = torch.arange(n_events) - torch.repeat_interleave(start, m)
        frac = (l.to(dt) / m[g_rep].to(dt))
        den = D[g_rep] - frac * E[g_rep]
        T = torch.zeros(J, dtype=dt).index_add_(0, g_rep, torch.log(den) + c)
        G = torch.zeros(J, dtype=dt).index_add_(0, g_rep, 1.0 / den)
        F = torch.zeros(J, dtype=dt).index_add_(0, g_rep, frac / den)
    pll = es[ds].sum() - T.sum()
    if reduction == "sum":
        norm = 1.0
    elif ties_method == "efron" and efron_mean_over == "event_times":
        norm = float(has.sum())
    else:
        norm = float(n_events)
    P = torch.cumsum(G, 0)
    g_sorted = dsf - w * (P[gid] - dsf * F[gid])
    grad = torch.empty(n, dtype=dt)
    grad[order] = -g_sorted / norm
    return -pll / norm, grad
