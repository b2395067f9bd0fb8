"""Labelled-row selection on B200.

``select_labelled(hazard, label, has_survival) -> (hazard_sel, time_sel, event_sel, n_events)`` replaces

    survival_mask = torch.tensor(has_survival, dtype=torch.bool, device=device)
    hazard_surv, time_surv, event_surv = hazard[survival_mask], label[survival_mask, 0], label[survival_mask, 1]

of scripts/training/partial_modality_training.py:401-406 (and simple_fusion.py:255-268) with one scan + scatter kernel
(csrc/compact.cu); gradients flow back to ``hazard`` through a scatter.  ONE device->host read of two counters sizes the
result (the reference synchronises three times here: ``survival_mask.sum() > 0``, the boolean index and
``event_surv.sum() > 0``); apply its skip rule with ``hazard_sel.shape[0] >= 2 and n_events > 0``.
"""
from __future__ import annotations

import torch

from . import _lib as L


class _Select(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hazard, label, keep):
        dev = hazard.device
        L.require_device(dev.index)
        lib = L.load()
        B = hazard.shape[0]
        hz = hazard.detach().reshape(-1).to(torch.float32).contiguous()
        lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
        kp = keep.to(device=dev, dtype=torch.bool).contiguous()
        oh = torch.empty(B, dtype=torch.float32, device=dev)
        ot = torch.empty(B, dtype=torch.float32, device=dev)
        oe = torch.empty(B, dtype=torch.bool, device=dev)
        oi = torch.empty(B, dtype=torch.int32, device=dev)
        cnt = torch.empty(2, dtype=torch.int64, device=dev)
        ws = torch.empty(lib.b200surv_compact_workspace_bytes(B), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.b200surv_compact_labelled(L.ptr(hz), L.ptr(lab), L.ptr(kp), B, L.ptr(oh), L.ptr(ot), L.ptr(oe),
                                                  L.ptr(oi), L.ptr(cnt), L.ptr(ws), ws.numel(), L.stream_ptr(dev)),
                    "b200surv_compact_labelled")
        n_sel, n_ev = (int(v) for v in cnt.cpu().tolist())     # the one synchronisation
        ctx.save_for_backward(oi[:n_sel])
        ctx.meta = (B, hazard.shape, hazard.dtype)
        ctx.mark_non_differentiable(ot, oe)
        ctx.n_events = n_ev
        return oh[:n_sel], ot[:n_sel], oe[:n_sel], torch.tensor(n_ev)

    @staticmethod
    def backward(ctx, g_h, _g_t, _g_e, _g_n):
        (idx,) = ctx.saved_tensors
        B, shape, dtype = ctx.meta
        dev = idx.device
        out = torch.empty(B, dtype=torch.float32, device=dev)
        g = None if g_h is None else g_h.detach().to(torch.float32).contiguous()
        n_sel = idx.numel()
        with torch.cuda.device(dev):
            L.check(L.load().b200surv_scatter_rows(L.ptr(g) if n_sel else None, L.ptr(idx) if n_sel else None, n_sel, B,
                                                   L.ptr(out), L.stream_ptr(dev)), "b200surv_scatter_rows")
        return out.reshape(shape).to(dtype), None, None


def select_labelled(hazard, label, has_survival):
    """hazard (B,) or (B,1) CUDA float; label (B,2) = (time, event 0/1); has_survival: bool tensor or sequence of B flags.
    Returns (hazard_sel, time_sel, event_sel [bool], n_events [int])."""
    if not hazard.is_cuda:
        raise L.B200SurvError("select_labelled has no CPU path: move the tensors to CUDA")
    keep = has_survival if isinstance(has_survival, torch.Tensor) else torch.tensor(list(has_survival), dtype=torch.bool)
    if keep.numel() != hazard.shape[0] or label.shape[0] != hazard.shape[0] or label.dim() != 2 or label.shape[1] != 2:
        raise ValueError("hazard (B,), label (B,2) and has_survival (B,) must agree")
    if hazard.shape[0] == 0:
        z = hazard.reshape(-1)
        return z, z.detach().clone(), torch.zeros(0, dtype=torch.bool, device=hazard.device), 0
    hs, ts, es, ne = _Select.apply(hazard, label, keep)
    return hs, ts, es, int(ne)
