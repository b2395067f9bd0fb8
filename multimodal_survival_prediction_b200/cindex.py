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
(simple_fusion.py:59-73: strict pairs only, no credit for risk ties, 0.5 when nothing is comparable).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L

COUNTER_NAMES = ("conc", "disc", "tied_risk", "conc_st", "disc_st", "tied_st")


def cindex_counts(estimate, event, time, tied_tol=1e-8, row_begin=0, row_end=None, algo=1, out=None):
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


def cindex_counts_shard(estimate, event, time, shard, n_shards, tied_tol=1e-8, out=None):
    """Six int64 counters of shard ``shard`` of ``n_shards`` (row tiles of the sorted event rows dealt out round-robin,
    b200surv_cindex_counts_shard); the shards' counters sum to cindex_counts(...) exactly."""
    dev = estimate.device
    L.require_device(dev.index)
    lib = L.load()
    n = estimate.numel()
    if out is None:
        out = torch.zeros(6, dtype=torch.int64, device=dev)
    if n == 0:
        return out
    wb = lib.b200surv_cindex_workspace_bytes(n, 1, 1)
    ws = torch.empty(max(wb, 256), dtype=torch.uint8, device=dev)
    rc = lib.b200surv_cindex_counts_shard(L.ptr(estimate), L.ptr(time), L.ptr(event), n, shard, n_shards,
                                          ctypes.c_float(tied_tol), L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "b200surv_cindex_counts_shard")
    return out


def cindex_counts_cohorts(estimate, event, time, offsets, tied_tol=1e-8, algo=1):
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
    raise ValueError("convention must be 'harrell' or 'fallback'")


class ConcordanceIndex:
    """``ConcordanceIndex(tied_tol=1e-8, checks=True)(estimate, event, time)`` -> 0-dim float32 tensor."""

    def __init__(self, tied_tol: float = 1e-8, checks: bool = True, *, convention: str = "harrell",
                 algo: int = 1):
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
        if est.numel() == 0:
            self.counts = [0] * 6
        else:
            with torch.cuda.device(dev):
                self.counts = cindex_counts(est, e, t, self.tied_tol, algo=self.algo).cpu().tolist()
        self.cindex = torch.tensor(cindex_from_counts(self.counts, self.convention), dtype=torch.float32,
                                   device=src_dev)
        return self.cindex
