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


class ValidationCohort:
    """Device-resident accumulation of a validation pass.

    The reference's ``validate`` (scripts/training/partial_modality_training.py:438-485) selects the labelled rows of
    every batch, adds the batch's Cox loss when ``n >= 2 and events > 0``, copies the three selected vectors to the
    host (``all_hazards.extend(hazard_surv.cpu().numpy())`` ..., :470-472), rebuilds tensors from Python lists and
    calls the C-index on the CPU (:478-481).  Here the compaction kernel writes every batch's kept rows straight
    behind the previous ones in three preallocated device buffers (its output pointers are the buffers' tails, no
    copy), the loss total stays on the device, and ``finish()`` runs the C-index on the buffers: one 16-byte D->H
    read per batch (the two counters the skip rule needs) and one at the end.

        cohort = ValidationCohort(capacity=len(val_dataset), device=device)
        for batch in loader: cohort.add(hazard, label, has_survival, loss_fn=cox_loss)
        avg_loss, c_index = cohort.finish(calculate_cindex)
    """

    def __init__(self, capacity: int, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.B200SurvError("ValidationCohort has no CPU path: pass a CUDA device")
        L.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        self.device = dev
        self.capacity = int(capacity)
        self.hazard = torch.empty(self.capacity, dtype=torch.float32, device=dev)
        self.time = torch.empty(self.capacity, dtype=torch.float32, device=dev)
        self.event = torch.empty(self.capacity, dtype=torch.bool, device=dev)
        self._index = torch.empty(0, dtype=torch.int32, device=dev)
        self._counts = torch.empty(2, dtype=torch.int64, device=dev)
        self._ws = torch.empty(0, dtype=torch.uint8, device=dev)
        self.n = 0                      # rows accumulated
        self.num_batches = 0            # batches that contributed a loss (the reference's num_batches)
        self.total_loss = torch.zeros((), dtype=torch.float32, device=dev)

    def add(self, hazard, label, has_survival, loss_fn=None):
        """Append the labelled rows of one batch; returns the number of rows appended (0 when the reference's rule
        ``n >= 2 and event.sum() > 0`` skips the batch: like the reference, a skipped batch adds no rows either)."""
        lib = L.load()
        dev = self.device
        B = int(hazard.shape[0])
        if B == 0:
            return 0
        keep = has_survival if isinstance(has_survival, torch.Tensor) else torch.tensor(list(has_survival), dtype=torch.bool)
        if keep.numel() != B or label.dim() != 2 or label.shape[0] != B or label.shape[1] != 2:
            raise ValueError("hazard (B,), label (B,2) and has_survival (B,) must agree")
        if self.n + B > self.capacity:
            raise ValueError(f"ValidationCohort capacity {self.capacity} exceeded ({self.n} rows held, batch of {B})")
        hz = hazard.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
        kp = keep.to(device=dev, dtype=torch.bool).contiguous()
        need = lib.b200surv_compact_workspace_bytes(B)
        if self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if self._index.numel() < B:
            self._index = torch.empty(B, dtype=torch.int32, device=dev)
        off = self.n
        with torch.cuda.device(dev):
            L.check(lib.b200surv_compact_labelled(
                L.ptr(hz), L.ptr(lab), L.ptr(kp), B, L.ptr(self.hazard[off:]), L.ptr(self.time[off:]),
                L.ptr(self.event[off:]), L.ptr(self._index), L.ptr(self._counts), L.ptr(self._ws), self._ws.numel(),
                L.stream_ptr(dev)), "b200surv_compact_labelled")
        n_sel, n_ev = (int(v) for v in self._counts.cpu().tolist())
        if n_sel < 2 or n_ev == 0:
            return 0                    # rows written past self.n are simply overwritten by the next batch
        if loss_fn is not None:
            with torch.no_grad():
                self.total_loss += loss_fn(self.hazard[off:off + n_sel], self.event[off:off + n_sel],
                                           self.time[off:off + n_sel]).to(torch.float32)
        self.num_batches += 1
        self.n += n_sel
        return n_sel

    def vectors(self):
        """(hazard, event [bool], time) of everything accumulated, views of the device buffers."""
        return self.hazard[:self.n], self.event[:self.n], self.time[:self.n]

    def finish(self, cindex_fn=None):
        """(avg_loss, c_index) with the reference's conventions: 0 loss without batches, 0.5 without rows (:474-483)."""
        avg = float(self.total_loss.item()) / self.num_batches if self.num_batches > 0 else 0
        if self.n == 0:
            return avg, 0.5
        if cindex_fn is None:
            from .cindex import ConcordanceIndex
            cindex_fn = lambda h, e, t: ConcordanceIndex()(h, e, t).item()  # noqa: E731
        c = cindex_fn(*self.vectors())
        return avg, float(c)
