"""Vectorised float64 Cox negative partial log-likelihood oracle: loss and gradient, O(n log n).

TEST INFRASTRUCTURE -- see oracle/__init__.py (parity with torchsurv is UNPINNED; the
published Breslow/Efron formulas are restated, validated against oracle/cox_def.py and,
on tie-free inputs, against the reference's runnable fallback loss
scripts/training/partial_modality_training.py:296-311 through tests/golden/).

Boundary mirrored: ``neg_partial_log_likelihood(log_hz, event, time, ties_method="efron",
reduction="mean")`` as called at partial_modality_training.py:285-288.

Derivation used by both this oracle and the CUDA kernels.  Let groups g = distinct times in
ascending order, w_j = exp(eta_j), S_g = sum of w over rows with time g, E_g = the same over
event rows, m_g = #event rows, D_g = sum_{g' >= g} S_g' (risk-set sum).  Then

    pll   = sum_{i: event} eta_i - sum_g T_g
    Efron : T_g = sum_{l<m_g} log(D_g - (l/m_g) E_g),  G_g = sum_l 1/(D_g - (l/m_g)E_g),
            F_g = sum_l (l/m_g)/(D_g - (l/m_g)E_g)
    Breslow: T_g = m_g log D_g, G_g = m_g / D_g, F_g = 0
    d pll / d eta_i = event_i - w_i * (P_{g(i)} - event_i * F_{g(i)}),   P_g = sum_{g' <= g} G_g'
    loss = -pll / normaliser
"""
from __future__ import annotations

import numpy as np

TIES = ("efron", "breslow")


def cox_nll(log_hz, event, time, ties_method="efron", reduction="mean",
            efron_mean_over="event_times", return_grad=True, dtype=np.float64):
    """Return (loss, grad) -- grad is d loss / d log_hz in the ORIGINAL row order.

    ``dtype`` is the arithmetic type (float64 for the oracle; float32 is used only when this
    routine is timed as the CPU baseline port of the reference's fp32 arithmetic).
    """
    if ties_method not in TIES:
        raise ValueError(f"ties_method {ties_method!r}")
    if reduction not in ("mean", "sum"):
        raise ValueError(f"reduction {reduction!r}")
    eta = np.asarray(log_hz).astype(dtype, copy=False)
    ev = np.asarray(event).astype(bool, copy=False)
    t = np.asarray(time)
    n = eta.shape[0]
    if n == 0 or not ev.any():
        return 0.0, np.zeros(n, dtype=dtype)

    order = np.argsort(t, kind="stable")
    ts, es, ds = t[order], eta[order], ev[order]
    head = np.empty(n, dtype=bool)
    head[0] = True
    np.not_equal(ts[1:], ts[:-1], out=head[1:])
    gid = np.cumsum(head) - 1
    J = int(gid[-1]) + 1

    c = es.max()
    w = np.exp(es - c)
    dsf = ds.astype(dtype)
    S = np.bincount(gid, weights=w, minlength=J).astype(dtype)
    E = np.bincount(gid, weights=w * dsf, minlength=J).astype(dtype)
    m = np.bincount(gid, weights=dsf, minlength=J).astype(np.int64)
    D = np.cumsum(S[::-1])[::-1]

    sum_ev_eta = float(np.sum(es[ds], dtype=np.float64))
    n_events = int(m.sum())
    has_ev = m > 0
    n_event_times = int(has_ev.sum())

    if ties_method == "breslow":
        mm = m.astype(dtype)
        T = np.where(has_ev, mm * (np.log(np.where(has_ev, D, 1.0)) + c), 0.0)
        G = np.where(has_ev, mm / np.where(has_ev, D, 1.0), 0.0)
        F = np.zeros(J, dtype=dtype)
    else:
        g_rep = np.repeat(np.arange(J), m)
        start = np.cumsum(m) - m
        l = np.arange(n_events) - np.repeat(start, m)
        frac = (l / m[g_rep]).astype(dtype)
        den = D[g_rep] - frac * E[g_rep]
        T = np.bincount(g_rep, weights=np.log(den) + c, minlength=J)
        G = np.bincount(g_rep, weights=1.0 / den, minlength=J)
        F = np.bincount(g_rep, weights=frac / den, minlength=J)

    pll = sum_ev_eta - float(np.sum(T, dtype=np.float64))
    if reduction == "sum":
        norm = 1.0
    elif ties_method == "efron" and efron_mean_over == "event_times":
        norm = float(n_event_times)
    else:
        norm = float(n_events)
    loss = -pll / norm
    if not return_grad:
        return loss, None
    P = np.cumsum(G)
    g_sorted = dsf - w * (P[gid] - dsf * F[gid])
    grad = np.empty(n, dtype=dtype)
    grad[order] = (-g_sorted / norm).astype(dtype)
    return loss, grad


def cox_nll_segmented(log_hz, event, time, seg_offsets, **kw):
    """Independent cohorts packed back to back (BASELINE.json configs[4] CV sweep)."""
    losses, grads = [], []
    for s in range(len(seg_offsets) - 1):
        a, b = int(seg_offsets[s]), int(seg_offsets[s + 1])
        l, g = cox_nll(log_hz[a:b], event[a:b], time[a:b], **kw)
        losses.append(l)
        grads.append(g)
    return np.asarray(losses), (np.concatenate(grads) if grads else np.zeros(0))
