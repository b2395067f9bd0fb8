"""Harrell's concordance index on B200 -- host-side mirror of torchsurv's ``ConcordanceIndex``.

The reference builds a fresh object per call and passes CPU tensors:
``ConcordanceIndex()(hazard, event.bool(), time).item()`` (scripts/training/
partial_modality_training.py:290-294, simple_fusion.py:330-331).  The pair counting runs in
libb200surv.so (csrc/cindex.cu); CPU inputs are copied to the current CUDA device; there is no CPU
path.  The result is a 0-dim float32 tensor on the input's device (torchsurv's return dtype,
SURVEY.md section 6); the exact int64 counters stay available on the object.

Unverifiable torchsurv conventions are explicit: ``convention="harrell"`` is
(C + T/2)/(C + D + T) over strict pairs plus same-time event-vs-censored pairs with
``tied_tol=1e-8`` on |estimate difference|; ``convention="fallback"`` is the reference's in-repo rule
(simple_fusion.py:59-73: strict pairs only, no credit for risk ties, 0.5 when nothing is comparable);
``convention="lifelines"`` is the rule of ``lifelines.utils.concordance_index`` -- the reference's other
fallback, ``concordance_index(time, -hazard, event)`` at partial_modality_training.py:313-319 and
scripts/analysis/evaluate_model.py:41-45: the same comparable pairs as "harrell" (an event row against every
later row and against the censored rows of its own time), risk ties by EXACT equality (``tied_tol`` is ignored:
the counters are taken with tolerance 0), float64 result, ZeroDivisionError when nothing is comparable.
``concordance_index_lifelines`` below keeps that function's signature (``shim/lifelines``).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L

COUNTER_NAMES = ("conc", "disc", "tied_risk", "conc_st", "disc_st", "tied_st")


def cindex_counts(estimate, event, time, tied_tol=1e-8, row_begin=0, row_end=None, algo=2, out=None):
    """int64[6] pair counters (device tensor) for rows [row_begin,row_end) x all columns.
    Asynchronous; ADDS into ``out`` if given."""
    dev = estimate.device
    L.require_device(dev.index)
    lib = L.load()
    n = estimate.numel()
    row_end = n if row_end is None else row_end
    if out is None:
        out = torch.zeros(6, dtype=torch.int64, device=dev)
    wb = lib.b200surv_cindex_workspace_bytes(n, 1, algo)
    ws = torch.empty(max(wb, 256), dtype=torch.uint8, device=dev)
    rc = lib.b200surv_cindex_counts(L.ptr(estimate), L.ptr(time), L.ptr(event), None, n, 1, row_begin, row_end,
                                    ctypes.c_float(tied_tol), algo, L.ptr(out), L.ptr(ws), ws.numel(),
                                    L.stream_ptr(dev))
    L.check(rc, "b200surv_cindex_counts")
    return out


def cindex_counts_shard(estimate, event, time, shard, n_shards, tied_tol=1e-8, out=None, algo=1):
    """Six int64 counters of shard ``shard`` of ``n_shards`` (row tiles of the sorted event rows dealt out round-robin,
    b200surv_cindex_counts_shard / _shard_algo); the shards' counters sum to cindex_counts(...) exactly."""
    dev = estimate.device
    L.require_device(dev.index)
    lib = L.load()
    n = estimate.numel()
    if out is None:
        out = torch.zeros(6, dtype=torch.int64, device=dev)
    if n == 0:
        return out
    wb = lib.b200surv_cindex_workspace_bytes(n, 1, algo)
    ws = torch.empty(max(wb, 256), dtype=torch.uint8, device=dev)
    if algo == 1:
        rc = lib.b200surv_cindex_counts_shard(L.ptr(estimate), L.ptr(time), L.ptr(event), n, shard, n_shards,
                                              ctypes.c_float(tied_tol), L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    else:
        rc = lib.b200surv_cindex_counts_shard_algo(L.ptr(estimate), L.ptr(time), L.ptr(event), n, shard, n_shards,
                                                   ctypes.c_float(tied_tol), algo, L.ptr(out), L.ptr(ws), ws.numel(),
                                                   L.stream_ptr(dev))
    L.check(rc, "b200surv_cindex_counts_shard")
    return out


def cindex_counts_cohorts(estimate, event, time, offsets, tied_tol=1e-8, algo=2):
    """int64[n_cohorts][6] pair counters for cohorts packed back to back; ``offsets`` is a host sequence of
    n_cohorts+1 row offsets (the CV sweep: one C-index per fold and replica).  Asynchronous."""
    dev = estimate.device
    L.require_device(dev.index)
    lib = L.load()
    offs = [int(o) for o in offsets]
    nc = len(offs) - 1
    if nc < 1 or offs[0] != 0 or offs[-1] != estimate.numel() or any(b < a for a, b in zip(offs, offs[1:])):
        raise ValueError("offsets must run from 0 to n, non-decreasing, with at least one cohort")
    out = torch.zeros(nc, 6, dtype=torch.int64, device=dev)
    n_max = max(b - a for a, b in zip(offs, offs[1:]))
    wb = (lib.b200surv_cindex_workspace_bytes(n_max, 1, algo) + 255) // 256 * 256
    ws = torch.empty(max(wb, 256) * min(8, nc), dtype=torch.uint8, device=dev)   # room for 8 cohorts in flight
    host = (ctypes.c_int64 * (nc + 1))(*offs)
    rc = lib.b200surv_cindex_counts_cohorts(L.ptr(estimate), L.ptr(time), L.ptr(event), host, nc,
                                            ctypes.c_float(tied_tol), algo, L.ptr(out), L.ptr(ws), ws.numel(),
                                            L.stream_ptr(dev))
    L.check(rc, "b200surv_cindex_counts_cohorts")
    return out


def cindex_from_counts(counts, convention="harrell"):
    """float64 ratio from the six counters (host ints)."""
    c = [int(x) for x in counts]
    if convention == "harrell":
        C, D, T = c[0] + c[3], c[1] + c[4], c[2] + c[5]
        den = C + D + T
        return (C + 0.5 * T) / den if den > 0 else 0.5
    if convention == "fallback":
        den = c[0] + c[1] + c[2]
        return c[0] / den if den > 0 else 0.5
    if convention == "lifelines":      # counters taken with tied_tol = 0; (correct + tied / 2) / pairs
        C, D, T = c[0] + c[3], c[1] + c[4], c[2] + c[5]
        den = C + D + T
        if den == 0:
            raise ZeroDivisionError("No admissable pairs in the dataset.")
        return (C + T / 2) / den
    raise ValueError("convention must be 'harrell', 'fallback' or 'lifelines'")


class ConcordanceIndex:
    """``ConcordanceIndex(tied_tol=1e-8, checks=True)(estimate, event, time)`` -> 0-dim float32 tensor."""

    def __init__(self, tied_tol: float = 1e-8, checks: bool = True, *, convention: str = "harrell",
                 algo: int = 2):
        self.tied_tol = float(tied_tol)
        self.checks = checks
        self.convention = convention
        self.algo = algo
        self.counts = None      # last call's six int64 counters (host list)
        self.cindex = None

    def __call__(self, estimate, event, time, weight=None, tmax=None, instate=True):
        if weight is not None or tmax is not None:
            raise NotImplementedError("weight / tmax (IPCW, truncated C-index) are outside the reference's "
                                      "call pattern and not implemented")
        if self.checks:
            for name, t in (("estimate", estimate), ("event", event), ("time", time)):
                if not isinstance(t, torch.Tensor):
                    raise TypeError(f"Input '{name}' should be a tensor")
            if event.dtype != torch.bool:
                raise ValueError("Input 'event' should be of boolean type (use event.bool())")
            if estimate.dim() == 2 and estimate.shape[1] == 1:
                estimate = estimate[:, 0]
            if estimate.dim() != 1 or event.dim() != 1 or time.dim() != 1:
                raise ValueError("Inputs should be one-dimensional")
            if not (estimate.shape[0] == event.shape[0] == time.shape[0]):
                raise ValueError("Dimension mismatch between 'estimate', 'event' and 'time'")
        src_dev = estimate.device
        if estimate.is_cuda:
            dev = estimate.device
        else:
            if not torch.cuda.is_available():
                raise L.B200SurvError("no CUDA device: the B200 survival kernels have no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
        est = estimate.detach().to(device=dev, dtype=torch.float32).contiguous()
        t = time.detach().to(device=dev, dtype=torch.float32).contiguous()
        e = event.detach().to(device=dev).contiguous()
        tol = 0.0 if self.convention == "lifelines" else self.tied_tol
        if est.numel() == 0:
            self.counts = [0] * 6
        else:
            with torch.cuda.device(dev):
                counts = cindex_counts(est, e, t, tol, algo=self.algo)
                if self.checks:     # negative / NaN times: the check rides on the counters' device->host copy (one sync)
                    bad = torch.logical_not(t >= 0).any().to(torch.int64).reshape(1)
                    host = torch.cat([counts, bad]).cpu().tolist()
                    if host[6]:
                        raise ValueError("Input 'time' should be non-negative and free of NaN")
                    self.counts = host[:6]
                else:
                    self.counts = counts.cpu().tolist()
        # torchsurv returns float32; lifelines returns a float64 scalar
        out_dtype = torch.float64 if self.convention == "lifelines" else torch.float32
        self.cindex = torch.tensor(cindex_from_counts(self.counts, self.convention), dtype=out_dtype, device=src_dev)
        return self.cindex


def _as_f32_exact(x, dev, what):
    """Host array / Series / tensor -> fp32 device vector with the same order and equality pattern.

    lifelines compares in float64.  fp32 input (the reference hands over ``hazard.cpu().numpy()``,
    partial_modality_training.py:317) converts exactly.  float64 values that are not fp32-representable (risk scores
    read back from a CSV, evaluate_model.py:41-45) are replaced by their dense ranks, which keeps every ``<`` and
    ``==`` the pair counters evaluate (ranks are exact in fp32 up to 2^24 distinct values)."""
    import numpy as np
    if isinstance(x, torch.Tensor):
        v = x.detach()
    else:
        a = np.asarray(x)
        if a.dtype == object or a.dtype.kind not in "fiub":
            a = a.astype(np.float64)
        v = torch.from_numpy(np.ascontiguousarray(a))
    v = v.reshape(-1) if v.dim() == 2 and 1 in v.shape else v
    if v.dim() != 1:
        raise ValueError(f"'{what}' should be one-dimensional")
    v = v.to(dev)
    if v.dtype in (torch.float32, torch.float16, torch.bfloat16, torch.bool, torch.uint8, torch.int8, torch.int16):
        return v.to(torch.float32)
    v64 = v.to(torch.float64)
    if bool(torch.isnan(v64).any()):
        raise ValueError("NaNs detected in inputs, please correct or drop.")
    v32 = v64.to(torch.float32)
    if bool((v32.to(torch.float64) == v64).all()):
        return v32
    uniq, inv = torch.unique(v64, sorted=True, return_inverse=True)
    if uniq.numel() > (1 << 24):
        raise L.B200SurvError(f"'{what}': more than 2^24 distinct float64 values that are not fp32-representable")
    return inv.to(torch.float32)


def concordance_index_lifelines(event_times, predicted_scores, event_observed=None) -> float:
    """``lifelines.utils.concordance_index(event_times, predicted_scores, event_observed=None)`` on the B200.

    The reference's call is ``concordance_index(time, -hazard, event)`` (partial_modality_training.py:313-319,
    scripts/analysis/evaluate_model.py:41-45): ``predicted_scores`` are higher for LONGER survival, so the pair
    counters run on ``estimate = -predicted_scores`` (negation is exact) with tolerance 0.  Returns a Python
    float formed in float64, raises ``ZeroDivisionError`` when no pair is comparable and ``ValueError`` on NaNs or
    mismatched lengths, like lifelines.  Infinite scores are rejected (lifelines would compare them as equal;
    the fp32 pair predicate cannot)."""
    if not torch.cuda.is_available():
        raise L.B200SurvError("no CUDA device: the B200 survival kernels have no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    for x in (event_times, predicted_scores, event_observed):
        if isinstance(x, torch.Tensor) and x.is_cuda:
            dev = x.device
            break
    with torch.cuda.device(dev):
        t = _as_f32_exact(event_times, dev, "event_times")
        s = _as_f32_exact(predicted_scores, dev, "predicted_scores")
        if event_observed is None:
            e = torch.ones(t.numel(), dtype=torch.bool, device=dev)
        else:
            e = _as_f32_exact(event_observed, dev, "event_observed") != 0
        if not (t.numel() == s.numel() == e.numel()):
            raise ValueError("Observed events must be 1-dimensional of same length as event times")
        if bool(torch.isnan(t).any()) or bool(torch.isnan(s).any()):
            raise ValueError("NaNs detected in inputs, please correct or drop.")
        if bool(torch.isinf(s).any()):
            raise ValueError("infinite predicted_scores are not supported")
        counts = [0] * 6 if t.numel() == 0 else cindex_counts((-s).contiguous(), e.contiguous(), t.contiguous(),
                                                              0.0).cpu().tolist()
    return float(cindex_from_counts(counts, "lifelines"))
