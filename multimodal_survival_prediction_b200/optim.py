"""Fused gradient clipping + Adam / AdamW on B200 -- the end of the reference's training step.

``ClipAdam(params, lr, betas, eps, weight_decay, max_norm=1.0, adamw=False)`` replaces the pair
``torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); optimizer.step()`` of
scripts/training/partial_modality_training.py:427-428 (``optim.Adam(lr, weight_decay=1e-4)``, :536) and of
simple_fusion.py:273-274 (``optim.AdamW``, :391) with two passes over the gradients in libb200surv.so
(csrc/optim.cu, ``b200surv_clip_adam_step``).  Same update rule as torch.optim.Adam / AdamW (no amsgrad); ``step()``
returns the total gradient norm before clipping (a device scalar, like clip_grad_norm_).  fp32 CUDA parameters only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0, adamw=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm, adamw=adamw))
        self._ws = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        total_norm = None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise L.B200SurvError("ClipAdam has no CPU path: move the parameters to CUDA")
            L.require_device(dev.index)
            ms, vs, gs = [], [], []
            for p in ps:
                if p.dtype is not torch.float32 or not p.is_contiguous():
                    raise L.B200SurvError("ClipAdam handles contiguous fp32 parameters")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
                g = p.grad
                gs.append(g if (g.dtype is torch.float32 and g.is_contiguous()) else g.float().contiguous())
            group["step"] = group.get("step", 0) + 1
            for p in ps:
                self.state[p]["step"] = group["step"]
            n = len(ps)
            arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
            numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
            wb = lib.b200surv_clip_adam_workspace_bytes(numel, n)
            if self._ws is None or self._ws.numel() < wb or self._ws.device != dev:
                self._ws = torch.empty(wb, dtype=torch.uint8, device=dev)
            norm = torch.empty(1, dtype=torch.float32, device=dev)
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                rc = lib.b200surv_clip_adam_step(arr(ps), arr(gs), arr(ms), arr(vs), numel, n,
                                                 ctypes.c_float(group["max_norm"] or 0.0), ctypes.c_float(group["lr"]),
                                                 ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(group["eps"]),
                                                 ctypes.c_float(group["weight_decay"]), int(bool(group["adamw"])),
                                                 group["step"], L.ptr(norm), L.ptr(self._ws), self._ws.numel(),
                                                 L.stream_ptr(dev))
            L.check(rc, "b200surv_clip_adam_step")
            total_norm = norm.reshape(())
        self.last_total_norm = total_norm
        return loss
