"""Functional restatement of the reference's late-fusion survival heads.  TEST INFRASTRUCTURE ONLY.

Restates, on an explicit parameter dictionary that uses the reference's ``state_dict`` key names,
what these reference modules compute from the 128-d CT feature onward:

* gated head   -- PartialModalityNet.forward, scripts/training/partial_modality_training.py:234-277
  (layers declared at :193-232): rna 5005->512 (BatchNorm, ReLU, Dropout .3) ->128 (ReLU);
  clinical 1->32 (ReLU); per-modality mask multiply; gate 291->64 (ReLU)->3 softmax over
  [ct, rna, clin, mask]; gate-weighted concat (288); fusion 288->256 (BatchNorm, ReLU, Dropout .3)
  ->128 (ReLU); cox head 128->1.  Returns (hazard[B], gate[B,3]).
* ungated head -- MultiModalSurvivalNet.forward, scripts/training/final_multimodal.py:122-150:
  the same without mask and gate, returns hazard[B].

The arithmetic is plain PyTorch (which *is* the reference's arithmetic), so this oracle is pinned
bit-for-bit against the AST-extracted reference classes through tests/golden/head_*.npz
(oracle/gen_golden.py).  Dropout masks are explicit inputs (``drop1``/``drop2``: 0/1 keep masks)
so a kernel with its own RNG can be checked in train mode; ``None`` means no dropout.
"""
from __future__ import annotations

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
DROP_P = 0.3

HEAD_KEYS_COMMON = [
    ("rna_encoder.0.weight", lambda r: (512, r)), ("rna_encoder.0.bias", lambda r: (512,)),
    ("rna_encoder.1.weight", lambda r: (512,)), ("rna_encoder.1.bias", lambda r: (512,)),
    ("rna_encoder.1.running_mean", lambda r: (512,)), ("rna_encoder.1.running_var", lambda r: (512,)),
    ("rna_encoder.4.weight", lambda r: (128, 512)), ("rna_encoder.4.bias", lambda r: (128,)),
    ("clinical_encoder.0.weight", lambda r: (32, 1)), ("clinical_encoder.0.bias", lambda r: (32,)),
    ("fusion.0.weight", lambda r: (256, 288)), ("fusion.0.bias", lambda r: (256,)),
    ("fusion.1.weight", lambda r: (256,)), ("fusion.1.bias", lambda r: (256,)),
    ("fusion.1.running_mean", lambda r: (256,)), ("fusion.1.running_var", lambda r: (256,)),
    ("fusion.4.weight", lambda r: (128, 256)), ("fusion.4.bias", lambda r: (128,)),
    ("cox_head.weight", lambda r: (1, 128)), ("cox_head.bias", lambda r: (1,)),
]
HEAD_KEYS_GATE = [
    ("gate.0.weight", lambda r: (64, 291)), ("gate.0.bias", lambda r: (64,)),
    ("gate.2.weight", lambda r: (3, 64)), ("gate.2.bias", lambda r: (3,)),
]


def _bn(x, p, prefix, train, stats_out):
    g, b = p[prefix + ".weight"], p[prefix + ".bias"]
    if train:
        if x.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training")
        mu = x.mean(dim=0)
        var_b = x.var(dim=0, unbiased=False)
        if stats_out is not None:
            n = x.shape[0]
            stats_out[prefix] = (mu.detach(), (var_b * (n / (n - 1))).detach())
    else:
        mu, var_b = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    return (x - mu) / torch.sqrt(var_b + BN_EPS) * g + b


def _drop(x, keep):
    if keep is None:
        return x
    return x * keep.to(x.dtype) / (1.0 - DROP_P)


def _q_bf16(t):
    """Round to bf16 with a straight-through gradient (models the GEMM operand precision of the B200 path)."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def head_forward(p, ct_feat, rna, clinical, mask=None, train=False, drop1=None, drop2=None,
                 stats_out=None, bf16_operands=False):
    """p: dict of tensors keyed like the reference state_dict.  mask=None selects the ungated head.

    bf16_operands=True rounds both operands of the five large Linear layers to bf16 (products and sums stay in
    the working precision): the exact function the tensor-core path evaluates, so that ReLU/dropout decisions
    and therefore gradients can be compared tightly.  Returns (hazard, gate) for the gated head, hazard for the
    ungated one."""
    if bf16_operands:
        big = ("rna_encoder.0", "rna_encoder.4", "gate.0", "fusion.0", "fusion.4")

        def lin(x, w, b):
            if any(w is p.get(k + ".weight") for k in big):
                return torch.nn.functional.linear(_q_bf16(x), _q_bf16(w), b)
            return torch.nn.functional.linear(x, w, b)
    else:
        lin = torch.nn.functional.linear
    h = lin(rna, p["rna_encoder.0.weight"], p["rna_encoder.0.bias"])
    h = torch.relu(_bn(h, p, "rna_encoder.1", train, stats_out))
    h = _drop(h, drop1)
    rna_f = torch.relu(lin(h, p["rna_encoder.4.weight"], p["rna_encoder.4.bias"]))
    clin_f = torch.relu(lin(clinical, p["clinical_encoder.0.weight"], p["clinical_encoder.0.bias"]))
    ct_f = ct_feat
    gate = None
    if mask is not None:
        ct_f = ct_f * mask[:, 0:1]
        rna_f = rna_f * mask[:, 1:2]
        clin_f = clin_f * mask[:, 2:3]
        z = torch.cat([ct_f, rna_f, clin_f, mask], dim=1)
        z = torch.relu(lin(z, p["gate.0.weight"], p["gate.0.bias"]))
        gate = torch.softmax(lin(z, p["gate.2.weight"], p["gate.2.bias"]), dim=1)
        ct_f = ct_f * gate[:, 0:1]
        rna_f = rna_f * gate[:, 1:2]
        clin_f = clin_f * gate[:, 2:3]
    f = torch.cat([ct_f, rna_f, clin_f], dim=1)
    f = lin(f, p["fusion.0.weight"], p["fusion.0.bias"])
    f = torch.relu(_bn(f, p, "fusion.1", train, stats_out))
    f = _drop(f, drop2)
    f = torch.relu(lin(f, p["fusion.4.weight"], p["fusion.4.bias"]))
    hazard = lin(f, p["cox_head.weight"], p["cox_head.bias"]).squeeze(1)
    return (hazard, gate) if mask is not None else hazard


def bn_running_update(p, stats_out):
    """running <- (1 - momentum) * running + momentum * (batch mean, UNBIASED batch var)."""
    for prefix, (mu, var_u) in stats_out.items():
        p[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * p[prefix + ".running_mean"] + BN_MOMENTUM * mu
        p[prefix + ".running_var"] = (1 - BN_MOMENTUM) * p[prefix + ".running_var"] + BN_MOMENTUM * var_u


def gate_entropy_loss(gate, eps=1e-8):
    """partial_modality_training.py:322-331: -(mean over rows of the gate entropy)."""
    return -(-(gate * torch.log(gate + eps)).sum(dim=1)).mean()


def init_head_params(rna_dim=5005, gated=True, seed=0, dtype=torch.float32):
    """Deterministic parameters with nn.Linear-like scale (uniform +-1/sqrt(fan_in)), BN gamma in
    [0.5,1.5], beta/running_mean small, running_var in [0.5,1.5] -- a synthetic stand-in for a
    trained checkpoint (the reference ships none)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    keys = HEAD_KEYS_COMMON + (HEAD_KEYS_GATE if gated else [])
    for name, shp in keys:
        shape = shp(rna_dim)
        if name.endswith("running_var") or (name.endswith(".1.weight")):
            t = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean") or name.endswith(".1.bias"):
            t = 0.1 * (torch.rand(shape, generator=g) - 0.5)
        else:
            fan_in = shape[1] if len(shape) == 2 else {512: rna_dim, 128: 512, 32: 1, 64: 291, 3: 64,
                                                      256: 288, 1: 128}.get(shape[0], 128)
            bound = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        p[name] = t.to(dtype)
    return p
