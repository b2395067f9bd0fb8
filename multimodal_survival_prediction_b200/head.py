"""Late-fusion survival heads on B200 -- host-side mirror of the reference's model classes.

``PartialModalityNet(rna_dim=5005, clinical_dim=1).forward(ct, rna, clinical, mask) -> (hazard, gate)``
mirrors scripts/training/partial_modality_training.py:165-277, and
``MultiModalSurvivalNet(rna_dim=5005, clinical_dim=1).forward(ct, rna, clinical) -> hazard`` mirrors
scripts/training/final_multimodal.py:59-150: same constructor arguments, forward signatures, return
types and ``state_dict`` keys/shapes (so ``.pth`` files interchange).  The sub-modules exist as
parameter containers; everything from the 128-d CT feature onward runs in libb200surv.so
(csrc/head.cu + csrc/gemm_tc.cu: tcgen05/TMEM GEMMs, fused BatchNorm/dropout/gate kernels) through
``b200surv_head_fwd`` / ``b200surv_head_bwd``.  The CT encoder (the reference's CNN branch) is
``ctenc.CTEncoderCNN`` on the ``b200surv_ct_*`` primitives; the MONAI DenseNet121 branch is out of scope.  There is no
CPU path.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib as L

_P_FIELDS = ["rna0_w", "rna0_b", "bn1_w", "bn1_b", "bn1_rm", "bn1_rv", "rna4_w", "rna4_b", "clin_w", "clin_b",
             "gate0_w", "gate0_b", "gate2_w", "gate2_b", "fus0_w", "fus0_b", "bn2_w", "bn2_b", "bn2_rm", "bn2_rv",
             "fus4_w", "fus4_b", "cox_w", "cox_b"]
_G_FIELDS = [f for f in _P_FIELDS if not f.endswith(("_rm", "_rv"))]
# field -> reference state_dict key
KEYS = {"rna0_w": "rna_encoder.0.weight", "rna0_b": "rna_encoder.0.bias", "bn1_w": "rna_encoder.1.weight",
        "bn1_b": "rna_encoder.1.bias", "bn1_rm": "rna_encoder.1.running_mean", "bn1_rv": "rna_encoder.1.running_var",
        "rna4_w": "rna_encoder.4.weight", "rna4_b": "rna_encoder.4.bias", "clin_w": "clinical_encoder.0.weight",
        "clin_b": "clinical_encoder.0.bias", "gate0_w": "gate.0.weight", "gate0_b": "gate.0.bias",
        "gate2_w": "gate.2.weight", "gate2_b": "gate.2.bias", "fus0_w": "fusion.0.weight", "fus0_b": "fusion.0.bias",
        "bn2_w": "fusion.1.weight", "bn2_b": "fusion.1.bias", "bn2_rm": "fusion.1.running_mean",
        "bn2_rv": "fusion.1.running_var", "fus4_w": "fusion.4.weight", "fus4_b": "fusion.4.bias",
        "cox_w": "cox_head.weight", "cox_b": "cox_head.bias"}


class HeadParams(ctypes.Structure):
    _fields_ = [(f, ctypes.c_void_p) for f in _P_FIELDS]


class HeadGrads(ctypes.Structure):
    _fields_ = [(f, ctypes.c_void_p) for f in _G_FIELDS]


def _f32c(t, dev):
    # fast path: already what the C ABI wants (every torch call costs microseconds; the head makes ~50 of them)
    if t.dtype is torch.float32 and t.device == dev and t.is_contiguous():
        return t.detach() if t.requires_grad else t
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _HeadFn(torch.autograd.Function):
    """(ct_feat, rna, clinical, mask|None, training, dropout_p, seed, want_masks, staged_saved|None, *24 parameter tensors).

    staged_saved: a caller-owned ``saved`` buffer into which ``stage_rna`` has already written the bf16 copy of ``rna``
    (B200SURV_HEAD_X_STAGED): the forward pass then does not read ``rna`` (CUDA-graph replays, GraphedHeadStep)."""

    @staticmethod
    def forward(ctx, ct_feat, rna, clinical, mask, training, dropout_p, seed, want_masks, staged_saved, *params):
        dev = ct_feat.device
        if dev.type != "cuda":
            raise L.B200SurvError("the B200 fusion head has no CPU path: move the module and its inputs to CUDA")
        L.require_device(dev.index)
        lib = L.load()
        gated = mask is not None
        B, rna_dim = rna.shape[0], rna.shape[1]
        ct_c, rna_c, clin_c = _f32c(ct_feat, dev), _f32c(rna, dev), _f32c(clinical, dev).reshape(B, 1)
        mask_c = _f32c(mask, dev) if gated else None
        tens = {}
        for f, t in zip(_P_FIELDS, params):
            tens[f] = None if t is None else (t.detach() if f.endswith(("_rm", "_rv")) else _f32c(t, dev))
        ps = HeadParams(**{f: (t.data_ptr() if t is not None else None) for f, t in tens.items()})
        sb = lib.b200surv_head_saved_bytes(B, rna_dim)
        wb = lib.b200surv_head_workspace_bytes(B, rna_dim)
        saved = torch.empty(sb, dtype=torch.uint8, device=dev) if staged_saved is None else staged_saved
        ws = torch.empty(wb, dtype=torch.uint8, device=dev)
        hazard = torch.empty(B, dtype=torch.float32, device=dev)
        gate = torch.empty(B, 3, dtype=torch.float32, device=dev) if gated else None
        keep1 = torch.empty(B, 512, dtype=torch.uint8, device=dev) if want_masks else None
        keep2 = torch.empty(B, 256, dtype=torch.uint8, device=dev) if want_masks else None
        if isinstance(seed, torch.Tensor):   # device-resident seed (CUDA-graph replays): pass its address
            ctx.seed_keep = seed
            training, seed = (L.HEAD_TRAIN_SEED_DEV if training else 0), seed.data_ptr()
            if training and staged_saved is not None:   # graphed steps: the forward pass itself advances the seed
                training |= L.HEAD_SEED_ADVANCE
        if staged_saved is not None:
            training = int(training) | L.HEAD_X_STAGED
        with torch.cuda.device(dev):
            rc = lib.b200surv_head_fwd(ctypes.byref(ps), L.ptr(ct_c), L.ptr(rna_c), L.ptr(clin_c), L.ptr(mask_c), B,
                                       rna_dim, int(training), ctypes.c_float(dropout_p), ctypes.c_uint64(seed),
                                       L.ptr(hazard), L.ptr(gate), L.ptr(keep1), L.ptr(keep2), L.ptr(saved), sb,
                                       L.ptr(ws), wb, L.stream_ptr(dev))
        L.check(rc, "b200surv_head_fwd")
        ctx.held = (tens, clin_c, mask_c, saved, ws)
        ctx.meta = (B, rna_dim, int(training), float(dropout_p), int(seed), gated, [None if p is None else (p.dtype, p.shape) for p in params])
        ctx.keep_masks = (keep1, keep2)
        outs = (hazard, gate) if gated else (hazard,)
        if want_masks:
            ctx.mark_non_differentiable(keep1, keep2)
            outs = outs + (keep1, keep2)
        return outs

    @staticmethod
    def backward(ctx, d_hazard, *rest):
        tens, clin_c, mask_c, saved, ws = ctx.held
        B, rna_dim, training, dropout_p, seed, gated, pmeta = ctx.meta
        dev = saved.device
        lib = L.load()
        d_gate = rest[0] if gated and len(rest) > 0 else None
        dh = torch.zeros(B, dtype=torch.float32, device=dev) if d_hazard is None else _f32c(d_hazard, dev)
        dg = None if d_gate is None else _f32c(d_gate, dev)
        # one allocation for all parameter gradients, handed out as views
        live = [f for f in _G_FIELDS if tens[f] is not None]
        flat = torch.empty(sum(tens[f].numel() for f in live), dtype=torch.float32, device=dev)
        grads = {f: None for f in _G_FIELDS}
        for f, g in zip(live, flat.split([tens[f].numel() for f in live])):
            grads[f] = g
        ps = HeadParams(**{f: (t.data_ptr() if t is not None else None) for f, t in tens.items()})
        gs = HeadGrads(**{f: (t.data_ptr() if t is not None else None) for f, t in grads.items()})
        d_ct = torch.empty(B, 128, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.b200surv_head_bwd(ctypes.byref(ps), ctypes.byref(gs), L.ptr(dh), L.ptr(dg), L.ptr(clin_c),
                                       L.ptr(mask_c), B, rna_dim, training, ctypes.c_float(dropout_p),
                                       ctypes.c_uint64(seed), L.ptr(d_ct), L.ptr(saved), saved.numel(), L.ptr(ws),
                                       ws.numel(), L.stream_ptr(dev))
        L.check(rc, "b200surv_head_bwd")
        out = [d_ct, None, None, None, None, None, None, None, None]
        for f, meta in zip(_P_FIELDS, pmeta):
            if f.endswith(("_rm", "_rv")) or meta is None:
                out.append(None)
            else:
                g = grads[f].view(meta[1])
                out.append(g if meta[0] is torch.float32 else g.to(meta[0]))
        return tuple(out)


def _param_list(module):
    """The 24 tensors of _P_FIELDS by direct attribute access (cheaper than walking named_parameters())."""
    r, f = module.rna_encoder, module.fusion
    g = getattr(module, "gate", None)
    return [r[0].weight, r[0].bias, r[1].weight, r[1].bias, r[1].running_mean, r[1].running_var, r[4].weight, r[4].bias,
            module.clinical_encoder[0].weight, module.clinical_encoder[0].bias,
            g[0].weight if g is not None else None, g[0].bias if g is not None else None,
            g[2].weight if g is not None else None, g[2].bias if g is not None else None,
            f[0].weight, f[0].bias, f[1].weight, f[1].bias, f[1].running_mean, f[1].running_var, f[4].weight, f[4].bias,
            module.cox_head.weight, module.cox_head.bias]


def head_saved_buffer(batch: int, rna_dim: int, device):
    """A ``saved`` buffer for ``stage_rna`` / ``fused_head(..., staged_saved=...)``."""
    return torch.empty(L.load().b200surv_head_saved_bytes(batch, rna_dim), dtype=torch.uint8, device=device)


def stage_rna(rna, saved):
    """bf16 copy of the batch ``rna`` (B, rna_dim) straight into ``saved`` (b200surv_head_stage_rna): what the forward pass
    would do first, done by the caller instead so that a captured graph needs no static fp32 copy of the batch."""
    dev = saved.device
    x = _f32c(rna, dev)
    with torch.cuda.device(dev):
        L.check(L.load().b200surv_head_stage_rna(L.ptr(x), x.shape[0], x.shape[1], L.ptr(saved), saved.numel(), L.stream_ptr(dev)),
                "b200surv_head_stage_rna")


def stage_batch(rna, saved, copies):
    """``stage_rna`` plus up to three fp32 copies ``(dst, src)`` into static buffers, all in ONE launch
    (b200surv_head_stage_batch)."""
    dev = saved.device
    x = _f32c(rna, dev)
    pairs = [(d, _f32c(s, dev)) for d, s in copies]
    n = len(pairs)
    src = (ctypes.c_void_p * 3)(*[p[1].data_ptr() for p in pairs], *([0] * (3 - n)))
    dst = (ctypes.c_void_p * 3)(*[p[0].data_ptr() for p in pairs], *([0] * (3 - n)))
    cnt = (ctypes.c_int64 * 3)(*[p[0].numel() for p in pairs], *([0] * (3 - n)))
    for d, s in pairs:
        if d.numel() != s.numel() or not d.is_contiguous() or d.dtype != torch.float32:
            raise ValueError("stage_batch: static buffers must be contiguous fp32 of the sources' sizes")
    with torch.cuda.device(dev):
        L.check(L.load().b200surv_head_stage_batch(L.ptr(x), x.shape[0], x.shape[1], L.ptr(saved), saved.numel(), src, dst, cnt, n,
                                                   L.stream_ptr(dev)), "b200surv_head_stage_batch")


def fused_head(module, ct_feat, rna, clinical, mask=None, want_masks=False, seed=None, staged_saved=None):
    """Run the head of ``module`` (a PartialModalityNet / MultiModalSurvivalNet from this file) on CUDA."""
    params = _param_list(module)
    training = module.training
    p_drop = float(module.rna_encoder[3].p) if training else 0.0
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) & ((1 << 62) - 1) if (training and p_drop > 0) else 0
    outs = _HeadFn.apply(ct_feat, rna, clinical, mask, training, p_drop, seed, want_masks, staged_saved, *params)
    if training:
        with torch.no_grad():
            torch._foreach_add_([module.rna_encoder[1].num_batches_tracked, module.fusion[1].num_batches_tracked], 1)
    return outs


def _ct_cnn():
    # the reference's non-MONAI CT encoder (partial_modality_training.py:179-190) on the b200surv_ct_* primitives:
    # same sub-modules and state_dict keys as the reference's nn.Sequential (ctenc.py)
    from .ctenc import CTEncoderCNN
    return CTEncoderCNN()


class _HeadBase(nn.Module):
    def __init__(self, rna_dim=5005, clinical_dim=1, gated=True):
        super().__init__()
        if clinical_dim != 1:
            raise NotImplementedError("the reference only uses clinical_dim=1 (age/100)")
        self.ct_encoder = _ct_cnn()
        self.use_monai = False
        self.ct_pool = nn.AdaptiveAvgPool3d(1)
        self.rna_encoder = nn.Sequential(nn.Linear(rna_dim, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.3),
                                         nn.Linear(512, 128), nn.ReLU())
        self.clinical_encoder = nn.Sequential(nn.Linear(clinical_dim, 32), nn.ReLU())
        if gated:
            self.gate = nn.Sequential(nn.Linear(128 + 128 + 32 + 3, 64), nn.ReLU(), nn.Linear(64, 3), nn.Softmax(dim=1))
        self.fusion = nn.Sequential(nn.Linear(128 + 128 + 32, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.3),
                                    nn.Linear(256, 128), nn.ReLU())
        self.cox_head = nn.Linear(128, 1)

    def _ct_features(self, ct):
        return self.ct_encoder(ct).view(ct.size(0), -1)


class PartialModalityNet(_HeadBase):
    """Gated head: forward(ct, rna, clinical, mask) -> (hazard [B], gate_weights [B,3]).

    ``skip_missing_ct`` (default False = the reference): the reference runs the CT encoder on every
    row, zero volumes of patients without imaging included, and multiplies their features by mask[:, 0] = 0 afterwards
    (partial_modality_training.py:245-259).  With the flag set the encoder only sees the rows that have a volume -- 142 of
    608 patients in the reference cohort -- and the others get zero features directly.  Hazards of an eval-mode model are
    unchanged; in training mode BatchNorm3d then takes its batch statistics over the present volumes only, which is NOT
    what the reference computes (contract a4), hence a flag and not the default.  Costs one device->host read (the number
    of present rows)."""

    def __init__(self, rna_dim=5005, clinical_dim=1, skip_missing_ct: bool = False):
        super().__init__(rna_dim, clinical_dim, gated=True)
        self.skip_missing_ct = skip_missing_ct

    def _ct_features_present(self, ct, mask):
        present = torch.nonzero(mask[:, 0] != 0).squeeze(1)          # (synchronises: the subset's size shapes the launch)
        feat = torch.zeros(ct.size(0), 128, dtype=ct.dtype, device=ct.device)
        if present.numel() == 0:
            return feat
        if present.numel() == ct.size(0):
            return self._ct_features(ct)
        return feat.index_copy(0, present, self._ct_features(ct.index_select(0, present)))

    def forward(self, ct, rna, clinical, mask):
        ct_feat = self._ct_features_present(ct, mask) if self.skip_missing_ct else self._ct_features(ct)
        hazard, gate = fused_head(self, ct_feat, rna, clinical, mask)
        return hazard, gate

    def forward_features(self, ct_feat, rna, clinical, mask):
        """Head only, from a precomputed CT feature [B,128] (benchmarks, BASELINE.json configs[1])."""
        return fused_head(self, ct_feat, rna, clinical, mask)


class MultiModalSurvivalNet(_HeadBase):
    """Ungated head: forward(ct, rna, clinical) -> hazard [B]."""

    def __init__(self, rna_dim=5005, clinical_dim=1):
        super().__init__(rna_dim, clinical_dim, gated=False)

    def forward(self, ct, rna, clinical):
        return fused_head(self, self._ct_features(ct), rna, clinical, None)[0]

    def forward_features(self, ct_feat, rna, clinical):
        return fused_head(self, ct_feat, rna, clinical, None)[0]


class GraphedHeadStep:
    """One training step of the head -- forward_features, ``loss_fn(*outputs)``, backward -- captured in ONE CUDA graph.

    The eager step is bound by the host (about a hundred kernel launches plus the autograd bookkeeping cost ~1 ms at
    B = 4096, three times the GPU work); a replay costs one launch.  Shapes are fixed at construction.  ``step(...)``
    copies a batch into the static input buffers -- the RNA matrix is converted to bf16 straight into the graph's saved
    buffer instead (``stage_rna``; to drive ``replay()`` yourself, write ct / clinical / mask into ``self.inputs[0 / 2 / 3]``
    and call ``stage_rna(rna, self.saved)``) --, replays, and returns ``(loss, outputs)``; the gradients are left in ``param.grad`` (static tensors that every replay
    overwrites -- step the optimizer before the next replay and do not set them to None).  The dropout seed lives in
    device memory and advances inside the graph, so every replay draws new masks (B200SURV_HEAD_TRAIN_SEED_DEV).
    """

    def __init__(self, module, ct_feat, rna, clinical, mask, loss_fn, warmup: int = 3):
        dev = ct_feat.device
        self.module, self.loss_fn = module, loss_fn
        self.inputs = [None if t is None else t.detach().clone() for t in (ct_feat, rna, clinical, mask)]
        # the RNA batch (82 MB at B = 4096) is not copied into a static fp32 buffer: step() converts every new batch to bf16
        # straight into the graph's saved buffer (stage_rna) and the captured forward starts from there
        self.saved = head_saved_buffer(rna.shape[0], rna.shape[1], dev)
        stage_rna(rna, self.saved)
        self.seed = torch.empty((), dtype=torch.int64, device=dev).random_()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):       # warm-up off the capture stream (lazy initialisation, allocator pools)
            for _ in range(warmup):
                self._eager()
        cur.wait_stream(side)
        for p in module.parameters():
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.outputs = self._eager()

    def _eager(self):
        # (the device-resident seed is advanced by the forward pass: B200SURV_HEAD_SEED_ADVANCE)
        outs = fused_head(self.module, *self.inputs, seed=self.seed, staged_saved=self.saved)
        loss = self.loss_fn(*outs)
        loss.backward()
        return loss, outs

    def replay(self):
        self.graph.replay()
        return self.loss, self.outputs

    def step(self, ct_feat, rna, clinical, mask=None):
        copies, big = [], []
        for i, (dst, src) in enumerate(zip(self.inputs, (ct_feat, rna, clinical, mask))):
            if dst is None or src is dst or i == 1:
                continue
            # small device-resident fp32 inputs ride along with the RNA cast; anything else (host tensors, CT volumes) is copied
            if src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and dst.numel() <= (1 << 22):
                copies.append((dst, src))
            else:
                big.append((dst, src))
        for dst, src in big:
            dst.copy_(src, non_blocking=True)
        if rna is not self.inputs[1]:
            # fp32 -> bf16 into the saved buffer; self.inputs[1] is not read by the graph
            stage_batch(rna, self.saved, copies)
        else:
            for dst, src in copies:
                dst.copy_(src, non_blocking=True)
        return self.replay()


class GraphedModelStep(GraphedHeadStep):
    """One training step of the WHOLE net from CT volumes -- CT encoder (ctenc.CTEncoderCNN), head, ``loss_fn(*outputs)``,
    backward -- captured in ONE CUDA graph: the reference's step at batch 4 (partial_modality_training.py:382-435,
    simple_fusion.py:255-275) is bound by the host when issued call by call (~200 launches).  Same contract as
    GraphedHeadStep with ``ct`` = the (B,1,D,H,W) volumes; ``loss_fn`` must not synchronise (the Cox loss does not for
    cohorts of <= 2048 rows; pass ``checks=False``)."""

    def _eager(self):
        ct_feat = self.module._ct_features(self.inputs[0])
        outs = fused_head(self.module, ct_feat, *self.inputs[1:], seed=self.seed, staged_saved=self.saved)
        loss = self.loss_fn(*outs)
        loss.backward()
        return loss, outs


class _GateEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, eps):
        dev = gate.device
        lib = L.load()
        g = _f32c(gate, dev)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.b200surv_gate_entropy_fwd(L.ptr(g), g.shape[0], ctypes.c_float(eps), L.ptr(out), L.stream_ptr(dev)),
                    "b200surv_gate_entropy_fwd")
        ctx.save_for_backward(g)
        ctx.meta = (float(eps), gate.dtype)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        eps, dtype = ctx.meta
        dev = g.device
        go = _f32c(grad_out.reshape(1), dev)
        dg = torch.empty_like(g)
        with torch.cuda.device(dev):
            L.check(L.load().b200surv_gate_entropy_bwd(L.ptr(g), L.ptr(go), g.shape[0], ctypes.c_float(eps), L.ptr(dg),
                                                       L.stream_ptr(dev)), "b200surv_gate_entropy_bwd")
        return (dg if dtype is torch.float32 else dg.to(dtype)), None


def gate_entropy_loss(gate_weights, eps=1e-8):
    """partial_modality_training.py:322-331: mean over the batch of sum_k g log(g + eps) (the negative gate entropy;
    its gradient enters the head through d_gate).  One fused kernel each way for the head's (B, 3) gate weights."""
    if not gate_weights.is_cuda:
        raise L.B200SurvError("gate_entropy_loss has no CPU path: the gate weights come from the CUDA head")
    if not (gate_weights.dim() == 2 and gate_weights.shape[1] == 3 and gate_weights.shape[0] >= 1):
        raise ValueError(f"gate_entropy_loss expects the head's (B, 3) gate weights, got {tuple(gate_weights.shape)} "
                         "(no eager path: the reference's gate has three modalities)")
    return _GateEntropy.apply(gate_weights, eps)
