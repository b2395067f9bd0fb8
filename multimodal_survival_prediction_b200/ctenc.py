"""CT encoder of the fusion nets on B200.

``CTEncoderCNN`` is the reference's non-MONAI CT branch, scripts/training/partial_modality_training.py:179-190
(identical in final_multimodal.py) -- ``nn.Sequential(Conv3d(1,32,3,2,1), BatchNorm3d, ReLU, Conv3d(32,64,3,2,1),
BatchNorm3d, ReLU, Conv3d(64,128,3,2,1), BatchNorm3d, ReLU, AdaptiveAvgPool3d(1))`` -- with the same sub-modules (so the
``state_dict`` keys ``0.weight`` ... ``7.num_batches_tracked`` and a reference ``.pth`` interchange) but a forward /
backward in ``b200surv_ct_encoder_fwd`` / ``_bwd`` (csrc/ctenc.cu: the ``b200surv_ct_*`` primitives and the tcgen05
GEMM, strung together in C so that a batch of 4 is not host-bound): channels-last activations,
first convolution direct, the other two as im2col + GEMM over chunks of samples whose patch matrix stays in L2,
BatchNorm3d as column statistics of the [rows][channels] matrix, bf16 GEMM operands with fp32 accumulation (the 2e-2
tolerance of the head).  Input (B, 1, D, H, W) float CUDA tensor, output (B, 128, 1, 1, 1) like the reference.
There is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib as L

class CtParams(ctypes.Structure):            # b200surv_ct_params (include/b200surv.h)
    _fields_ = [(n, ctypes.c_void_p * 3) for n in ("w", "b", "gamma", "beta", "run_mean", "run_var")]


class CtGrads(ctypes.Structure):             # b200surv_ct_grads
    _fields_ = [(n, ctypes.c_void_p * 3) for n in ("w", "b", "gamma", "beta")]


def _out(d):
    return (d - 1) // 2 + 1


def _params_struct(p, buffers):
    st = CtParams()
    for s in range(3):
        st.w[s], st.b[s], st.gamma[s], st.beta[s] = (t.data_ptr() for t in p[4 * s: 4 * s + 4])
        st.run_mean[s], st.run_var[s] = buffers[2 * s].data_ptr(), buffers[2 * s + 1].data_ptr()
    return st


class _CTEncFn(torch.autograd.Function):
    """b200surv_ct_encoder_fwd / _bwd: one C call each way (about 35 + 45 launches, no host work in between)."""

    @staticmethod
    def forward(ctx, ct, training, bn_buffers, *params):
        # params: (w, b, gamma, beta) x 3; bn_buffers: (running_mean, running_var) x 3, updated in place in training
        dev = ct.device
        L.require_device(dev.index)
        lib = L.load()
        B, _, D, H, W = ct.shape
        x = ct.detach().reshape(B, D, H, W).to(torch.float32).contiguous()
        p = [t.detach().to(torch.float32).contiguous() for t in params]
        bufs = list(bn_buffers)
        if any(t.dtype != torch.float32 or not t.is_contiguous() for t in bufs):
            raise L.B200SurvError("BatchNorm running statistics must be contiguous float32")
        saved = torch.empty(lib.b200surv_ct_encoder_saved_bytes(B, D, H, W), dtype=torch.uint8, device=dev)
        ws = torch.empty(lib.b200surv_ct_encoder_workspace_bytes(B, D, H, W), dtype=torch.uint8, device=dev)
        feat = torch.empty(B, 128, dtype=torch.float32, device=dev)
        pst = _params_struct(p, bufs)
        with torch.cuda.device(dev):
            L.check(lib.b200surv_ct_encoder_fwd(L.ptr(x), ctypes.byref(pst), B, D, H, W, int(training), L.ptr(feat), L.ptr(saved),
                                                saved.numel(), L.ptr(ws), ws.numel(), L.stream_ptr(dev)),
                    "b200surv_ct_encoder_fwd")
        ctx.keep = (x, p, bufs, saved, ws)
        ctx.dims = (B, D, H, W, bool(training))
        ctx.param_meta = [(t.shape, t.dtype) for t in params]
        return feat.to(ct.dtype).reshape(B, 128, 1, 1, 1)

    @staticmethod
    def backward(ctx, d_out):
        x, p, bufs, saved, ws = ctx.keep
        B, D, H, W, training = ctx.dims
        dev = x.device
        lib = L.load()
        dfeat = d_out.detach().reshape(B, 128).to(torch.float32).contiguous()
        g = [torch.empty(t.shape, dtype=torch.float32, device=dev) for t in p]
        gst = CtGrads()
        for s in range(3):
            gst.w[s], gst.b[s], gst.gamma[s], gst.beta[s] = (t.data_ptr() for t in g[4 * s: 4 * s + 4])
        pst = _params_struct(p, bufs)
        with torch.cuda.device(dev):
            L.check(lib.b200surv_ct_encoder_bwd(L.ptr(x), ctypes.byref(pst), L.ptr(dfeat), B, D, H, W, int(training),
                                                ctypes.byref(gst), L.ptr(saved), saved.numel(), L.ptr(ws), ws.numel(),
                                                L.stream_ptr(dev)), "b200surv_ct_encoder_bwd")
        grads = [t.reshape(shape).to(dtype) if ctx.needs_input_grad[3 + i] else None
                 for i, (t, (shape, dtype)) in enumerate(zip(g, ctx.param_meta))]
        return (None, None, None, *grads)


class CTEncoderCNN(nn.Sequential):
    """Drop-in for the reference's CNN ``ct_encoder`` (partial_modality_training.py:179-190): same sub-modules and
    ``state_dict``; ``forward(ct (B,1,D,H,W)) -> (B,128,1,1,1)`` runs on the B200 primitives.  The input needs no
    gradient in the reference (a CT volume) and gets none here."""

    def __init__(self):
        super().__init__(
            nn.Conv3d(1, 32, 3, stride=2, padding=1), nn.BatchNorm3d(32), nn.ReLU(),
            nn.Conv3d(32, 64, 3, stride=2, padding=1), nn.BatchNorm3d(64), nn.ReLU(),
            nn.Conv3d(64, 128, 3, stride=2, padding=1), nn.BatchNorm3d(128), nn.ReLU(),
            nn.AdaptiveAvgPool3d(1),
        )

    def forward(self, ct):
        if not ct.is_cuda:
            raise L.B200SurvError("CTEncoderCNN has no CPU path: move the module and its input to CUDA")
        if ct.dim() != 5 or ct.shape[1] != 1:
            raise ValueError("ct must be (B, 1, D, H, W)")
        convs, bns = (self[0], self[3], self[6]), (self[1], self[4], self[7])
        training = self.training
        if training and ct.shape[0] * _out(_out(_out(ct.shape[2]))) * _out(_out(_out(ct.shape[3]))) * _out(_out(_out(ct.shape[4]))) == 1:
            raise ValueError("Expected more than 1 value per channel when training")       # nn.BatchNorm3d's rule
        params, buffers = [], []
        for conv, bn in zip(convs, bns):
            params += [conv.weight, conv.bias, bn.weight, bn.bias]
            buffers += [bn.running_mean, bn.running_var]
        out = _CTEncFn.apply(ct, training, buffers, *params)
        if training:
            with torch.no_grad():
                for bn in bns:
                    bn.num_batches_tracked += 1
        return out
