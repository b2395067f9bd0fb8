"""CT encoder of the fusion nets on B200.

``CTEncoderCNN`` is the reference's non-MONAI CT branch, scripts/training/partial_modality_training.py:179-190
(identical in final_multimodal.py) -- ``nn.Sequential(Conv3d(1,32,3,2,1), BatchNorm3d, ReLU, Conv3d(32,64,3,2,1),
BatchNorm3d, ReLU, Conv3d(64,128,3,2,1), BatchNorm3d, ReLU, AdaptiveAvgPool3d(1))`` -- with the same sub-modules (so the
``state_dict`` keys ``0.weight`` ... ``7.num_batches_tracked`` and a reference ``.pth`` interchange) but a forward /
backward made of the ``b200surv_ct_*`` primitives (csrc/ctenc.cu) and the tcgen05 GEMM: channels-last activations,
first convolution direct, the other two as im2col + GEMM over chunks of samples whose patch matrix stays in L2,
BatchNorm3d as column statistics of the [rows][channels] matrix, bf16 GEMM operands with fp32 accumulation (the 2e-2
tolerance of the head).  Input (B, 1, D, H, W) float CUDA tensor, output (B, 128, 1, 1, 1) like the reference.
There is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L

COL_BYTES_IN_L2 = 64 << 20        # patch-matrix bytes per chunk of samples: half of the 126 MB L2


def _out(d):
    return (d - 1) // 2 + 1


def _gemm(lib, dev, a, lda, a_mn, b, ldb, b_mn, M, N, K, c=None, ldc=0, c_bf16=None, ldc_bf16=0, bias=None):
    L.check(lib.b200surv_gemm_bf16(L.ptr(a), lda, a_mn, L.ptr(b), ldb, b_mn, M, N, K, L.ptr(c), ldc, L.ptr(c_bf16),
                                   ldc_bf16, L.ptr(bias), 0, L.stream_ptr(dev)), "b200surv_gemm_bf16")


class _Stage:
    """Geometry of one stride-2 convolution stage."""

    def __init__(self, B, D, H, W, Cin, Cout):
        self.B, self.D, self.H, self.W, self.Cin, self.Cout = B, D, H, W, Cin, Cout
        self.Do, self.Ho, self.Wo = _out(D), _out(H), _out(W)
        self.vin, self.vox = D * H * W, self.Do * self.Ho * self.Wo
        self.R = B * self.vox
        self.K = 27 * Cin
        self.chunk = max(1, min(B, COL_BYTES_IN_L2 // (self.vox * self.K * 2)))
        self.nchunks = (B + self.chunk - 1) // self.chunk


class _CTEncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ct, training, bn_buffers, *params):
        # params: (w, b, gamma, beta) x 3; bn_buffers: (running_mean, running_var) x 3, updated in place in training
        dev = ct.device
        L.require_device(dev.index)
        lib = L.load()
        st = L.stream_ptr(dev)
        B, _, D, H, W = ct.shape
        x = ct.detach().reshape(B, D, H, W).to(torch.float32).contiguous()
        p = [t.detach().to(torch.float32).contiguous() for t in params]
        ws = torch.empty(lib.b200surv_ct_workspace_bytes(), dtype=torch.uint8, device=dev)
        chans = [1] + [int(p[4 * s].shape[0]) for s in range(3)]
        stages, d = [], (D, H, W)
        for s in range(3):
            sg = _Stage(B, d[0], d[1], d[2], chans[s], chans[s + 1])
            stages.append(sg)
            d = (sg.Do, sg.Ho, sg.Wo)
        hs, acts, mus, rstds, wrs = [], [], [], [], [None]
        a_prev = None
        with torch.cuda.device(dev):
            for s, sg in enumerate(stages):
                w, bias, gamma, beta = p[4 * s: 4 * s + 4]
                rm, rv = bn_buffers[2 * s], bn_buffers[2 * s + 1]
                h = torch.empty(sg.R, sg.Cout, dtype=torch.float32, device=dev)
                if s == 0:
                    L.check(lib.b200surv_ct_conv_first_fwd(L.ptr(x), L.ptr(w), L.ptr(bias), B, sg.D, sg.H, sg.W, sg.Cout,
                                                           L.ptr(h), st), "b200surv_ct_conv_first_fwd")
                else:
                    wr = torch.empty(sg.Cout, sg.K, dtype=torch.bfloat16, device=dev)
                    L.check(lib.b200surv_ct_weight_pack(L.ptr(w), sg.Cout, sg.Cin, L.ptr(wr), st), "b200surv_ct_weight_pack")
                    wrs.append(wr)
                    col = torch.empty(sg.chunk * sg.vox, sg.K, dtype=torch.bfloat16, device=dev)
                    for b0 in range(0, B, sg.chunk):
                        bc = min(sg.chunk, B - b0)
                        L.check(lib.b200surv_ct_im2col(L.ptr(a_prev[b0 * sg.vin:]), bc, sg.D, sg.H, sg.W, sg.Cin, L.ptr(col), st),
                                "b200surv_ct_im2col")
                        _gemm(lib, dev, col, sg.K, 0, wr, sg.K, 0, bc * sg.vox, sg.Cout, sg.K, c=h[b0 * sg.vox:], ldc=sg.Cout,
                              bias=bias)
                mu = torch.empty(sg.Cout, dtype=torch.float32, device=dev)
                rstd = torch.empty(sg.Cout, dtype=torch.float32, device=dev)
                L.check(lib.b200surv_ct_bn_stats(L.ptr(h), sg.R, sg.Cout, int(training), L.ptr(rm), L.ptr(rv), L.ptr(mu),
                                                 L.ptr(rstd), L.ptr(ws), ws.numel(), st), "b200surv_ct_bn_stats")
                if s < 2:
                    a = torch.empty(sg.R, sg.Cout, dtype=torch.bfloat16, device=dev)
                    L.check(lib.b200surv_ct_bn_relu(L.ptr(h), L.ptr(mu), L.ptr(rstd), L.ptr(gamma), L.ptr(beta), sg.R, sg.Cout,
                                                    L.ptr(a), st), "b200surv_ct_bn_relu")
                    acts.append(a)
                    a_prev = a
                else:
                    feat = torch.empty(B, sg.Cout, dtype=torch.float32, device=dev)
                    L.check(lib.b200surv_ct_bn_relu_pool(L.ptr(h), L.ptr(mu), L.ptr(rstd), L.ptr(gamma), L.ptr(beta), B, sg.vox,
                                                         sg.Cout, L.ptr(feat), st), "b200surv_ct_bn_relu_pool")
                hs.append(h); mus.append(mu); rstds.append(rstd)
        ctx.stages, ctx.training = stages, bool(training)
        ctx.saved = (x, p, hs, acts, mus, rstds, wrs, ws)
        ctx.out_dtype = ct.dtype
        ctx.param_meta = [(t.shape, t.dtype) for t in params]
        return feat.to(ct.dtype).reshape(B, stages[2].Cout, 1, 1, 1)

    @staticmethod
    def backward(ctx, d_out):
        x, p, hs, acts, mus, rstds, wrs, ws = ctx.saved
        stages, training = ctx.stages, int(ctx.training)
        dev = x.device
        lib = L.load()
        st = L.stream_ptr(dev)
        B = stages[0].B
        grads = [None] * 12
        with torch.cuda.device(dev):
            sg = stages[2]
            dfeat = d_out.detach().reshape(B, sg.Cout).to(torch.float32).contiguous()
            dA = torch.empty(sg.R, sg.Cout, dtype=torch.float32, device=dev)
            L.check(lib.b200surv_ct_pool_bwd(L.ptr(dfeat), B, sg.vox, sg.Cout, L.ptr(dA), st), "b200surv_ct_pool_bwd")
            for s in (2, 1, 0):
                sg = stages[s]
                w, bias, gamma, beta = p[4 * s: 4 * s + 4]
                dx = torch.empty(sg.R, sg.Cout, dtype=torch.bfloat16, device=dev)
                dgamma, dbeta, dbias = (torch.empty(sg.Cout, dtype=torch.float32, device=dev) for _ in range(3))
                L.check(lib.b200surv_ct_bn_bwd(L.ptr(hs[s]), L.ptr(dA), L.ptr(mus[s]), L.ptr(rstds[s]), L.ptr(gamma), L.ptr(beta),
                                               sg.R, sg.Cout, training, L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dbias),
                                               L.ptr(ws), ws.numel(), st), "b200surv_ct_bn_bwd")
                dw = torch.empty(sg.Cout, sg.Cin, 3, 3, 3, dtype=torch.float32, device=dev)
                if s == 0:
                    L.check(lib.b200surv_ct_conv_first_wgrad(L.ptr(x), L.ptr(dx), B, sg.D, sg.H, sg.W, sg.Cout, L.ptr(dw),
                                                             L.ptr(ws), ws.numel(), st), "b200surv_ct_conv_first_wgrad")
                else:
                    a_prev, wr = acts[s - 1], wrs[s]
                    col = torch.empty(sg.chunk * sg.vox, sg.K, dtype=torch.bfloat16, device=dev)
                    dcol = torch.empty(sg.chunk * sg.vox, sg.K, dtype=torch.bfloat16, device=dev)
                    nsl = [lib.b200surv_gemm_splitk_slices(sg.Cout, sg.K, min(sg.chunk, B - b0) * sg.vox)
                           for b0 in range(0, B, sg.chunk)]        # split-K slices every chunk's weight gradient writes
                    dwr = torch.empty(sum(nsl), sg.Cout, sg.K, dtype=torch.float32, device=dev)
                    sl0 = 0
                    dA_prev = torch.empty(B * sg.vin, sg.Cin, dtype=torch.float32, device=dev)
                    for k, b0 in enumerate(range(0, B, sg.chunk)):
                        bc = min(sg.chunk, B - b0)
                        rows = bc * sg.vox
                        dxc = dx[b0 * sg.vox:]
                        L.check(lib.b200surv_ct_im2col(L.ptr(a_prev[b0 * sg.vin:]), bc, sg.D, sg.H, sg.W, sg.Cin, L.ptr(col), st),
                                "b200surv_ct_im2col")
                        # dW (tap-major) = dx^T col : both operands read as stored (MN-major)
                        L.check(lib.b200surv_gemm_bf16_splitk(L.ptr(dxc), sg.Cout, 1, L.ptr(col), sg.K, 1, sg.Cout, sg.K, rows,
                                                              L.ptr(dwr[sl0:]), sg.K, st), "b200surv_gemm_bf16_splitk")
                        sl0 += nsl[k]
                        # dcol = dx W, then every input voxel gathers its taps
                        _gemm(lib, dev, dxc, sg.Cout, 0, wr, sg.K, 1, rows, sg.K, sg.Cout, c_bf16=dcol, ldc_bf16=sg.K)
                        L.check(lib.b200surv_ct_col2im(L.ptr(dcol), bc, sg.D, sg.H, sg.W, sg.Cin, L.ptr(dA_prev[b0 * sg.vin:]), st),
                                "b200surv_ct_col2im")
                    L.check(lib.b200surv_ct_weight_unpack(L.ptr(dwr), sum(nsl), sg.Cout, sg.Cin, L.ptr(dw), st),
                            "b200surv_ct_weight_unpack")
                    dA = dA_prev
                for j, g in enumerate((dw, dbias, dgamma, dbeta)):
                    shape, dtype = ctx.param_meta[4 * s + j]
                    grads[4 * s + j] = g.reshape(shape).to(dtype) if ctx.needs_input_grad[3 + 4 * s + j] else None
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
