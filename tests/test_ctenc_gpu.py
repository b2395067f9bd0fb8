"""CT encoder against the reference's nn.Sequential CNN branch,
scripts/training/partial_modality_training.py:179-190, run by PyTorch/cuDNN in fp32 on the same device."""
import copy

import pytest
import torch
from torch import nn

from multimodal_survival_prediction_b200.ctenc import CTEncoderCNN
from oracle.ctenc import matched_forward as _matched, reference_cnn as _reference_cnn

pytestmark = pytest.mark.gpu

OUT_TOL = 2e-2      # bf16 GEMM operands, fp32 accumulation (BASELINE.json: 2e-2 on the bf16 path)
GRAD_TOL = 3e-2     # per-tensor relative Frobenius error against the reference with bf16-rounded conv operands
GRAD_TOL_FP32 = 0.15  # ... and against the full-precision reference: in training mode BatchNorm's backward projects
#                       out most of dy, which amplifies the bf16 operand rounding to 5-9 % per tensor -- the matched
#                       reference shows the same 5-9 % against fp32 (scratch/ctenc_diag.py), eval mode stays < 1 %


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape,training", [((3, 1, 16, 16, 8), True), ((2, 1, 15, 13, 9), True), ((4, 1, 64, 64, 32), True),
                                            ((4, 1, 64, 64, 32), False), ((40, 1, 32, 32, 16), True)])
def test_ct_encoder_matches_torch_cnn(shape, training):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    torch.manual_seed(shape[0] * 1000 + shape[2])
    ref = _reference_cnn().to(dev)
    with torch.no_grad():                                   # non-trivial BatchNorm parameters and running statistics
        for m in ref:
            if isinstance(m, nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
    refm = copy.deepcopy(ref)
    ours = CTEncoderCNN().to(dev)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))   # a reference .pth loads as is
    ref.train(training); ours.train(training); refm.train(training)
    ct = torch.rand(shape, device=dev)
    ct[0].zero_()                                           # a patient without imaging: zero volume (:89)
    w = torch.randn(shape[0], 128, 1, 1, 1, device=dev)
    y_ref = ref(ct); (y_ref * w).sum().backward()
    (_matched(refm, ct) * w).sum().backward()
    y = ours(ct); (y * w).sum().backward()
    assert y.shape == y_ref.shape and y.dtype == y_ref.dtype
    assert float((y - y_ref).abs().max()) <= OUT_TOL * max(1.0, float(y_ref.abs().max())), "features"
    for (k, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
        if "running" in k:
            assert _rel(a, b) <= 1e-3, k
        if "num_batches" in k:
            assert int(a) == int(b), k
    live = [k for k, pb in ref.named_parameters() if not (training and k in ("0.bias", "3.bias", "6.bias"))]
    ga, gf, gm = (dict((k, p.grad) for k, p in m.named_parameters()) for m in (ours, ref, refm))
    errs = {k: (_rel(ga[k], gm[k]), _rel(ga[k], gf[k])) for k in live}
    assert max(e[0] for e in errs.values()) <= GRAD_TOL and max(e[1] for e in errs.values()) <= GRAD_TOL_FP32, errs
    if training:
        for i in (0, 3, 6):                                 # BatchNorm cancels the bias: exactly zero here
            assert float(ours[i].bias.grad.abs().max()) == 0.0


def test_ct_encoder_rejects_cpu_and_bad_shapes():
    enc = CTEncoderCNN()
    from multimodal_survival_prediction_b200 import B200SurvError
    with pytest.raises(B200SurvError):
        enc(torch.zeros(2, 1, 8, 8, 8))
    with pytest.raises(ValueError):
        enc.cuda()(torch.zeros(2, 2, 8, 8, 8, device="cuda"))


def test_full_model_step_with_ct_volumes():
    """PartialModalityNet.forward(ct, rna, clinical, mask) (partial_modality_training.py:234-277) with CT volumes:
    the gradient of a Cox loss reaches the CT encoder through the fused head; reference = the same head fed by the
    torch CNN (bf16-matched) with the same weights."""
    from multimodal_survival_prediction_b200 import head as ghead, neg_partial_log_likelihood, synth
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    B = 8
    m = ghead.PartialModalityNet().to(dev).train()
    m.rna_encoder[3].p = 0.0; m.fusion[3].p = 0.0            # dropout off: two passes must see the same network
    cnn = _reference_cnn().to(dev).train()
    cnn.load_state_dict(copy.deepcopy(m.ct_encoder.state_dict()))
    before = copy.deepcopy(m.state_dict())
    _, rna, clin, mask = [t.to(dev) for t in synth.modality_batch(B, seed=3)]
    ct = torch.rand(B, 1, 64, 64, 32, device=dev) * mask[:, 0].view(B, 1, 1, 1, 1)      # no imaging -> zero volume
    time = torch.arange(1, B + 1, device=dev).float()
    event = torch.tensor([1, 0, 1, 1, 0, 1, 0, 1], device=dev).bool()
    hz, gate = m(ct, rna, clin, mask)
    neg_partial_log_likelihood(hz, event, time).backward()
    g_ours = {k: p.grad.clone() for k, p in m.ct_encoder.named_parameters()}
    m.load_state_dict(before); m.zero_grad()
    hz2, _ = m.forward_features(_matched(cnn, ct).view(B, -1), rna, clin, mask)
    neg_partial_log_likelihood(hz2, event, time).backward()
    assert float((hz - hz2).abs().max()) <= OUT_TOL * max(1.0, float(hz2.abs().max()))
    errs = {k: _rel(g_ours[k], p.grad) for k, p in cnn.named_parameters() if k not in ("0.bias", "3.bias", "6.bias")}
    assert all(float(g.abs().max()) > 0 for k, g in g_ours.items() if k in errs)
    assert max(errs.values()) <= 2 * GRAD_TOL, errs           # two bf16 heads in front of the comparison


def test_ct_encoder_matches_reference_class_golden():
    """tests/golden/ct_encoder.npz: the reference's own PartialModalityNet (fp64, CPU) on 24 x 20 x 12 volumes -- CT
    features in eval mode, then one training step of the whole model and the gradients / running statistics of
    ct_encoder.* (oracle/gen_golden.py: gen_ct_encoder)."""
    import os
    import numpy as np
    from multimodal_survival_prediction_b200 import head as ghead
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ct_encoder.npz"))
    dev = torch.device("cuda", 0)
    m = ghead.PartialModalityNet(rna_dim=40).to(dev)
    sd0 = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0/")}
    m.load_state_dict(sd0)                                   # the reference's state_dict loads key for key
    ct, rna, clin, mask = (torch.from_numpy(g[k]).float().to(dev) for k in ("ct", "rna", "clinical", "mask"))
    m.eval()
    with torch.no_grad():
        feat = m.ct_encoder(ct).view(ct.shape[0], -1)
        hz = m(ct, rna, clin, mask)[0]
    ref = torch.from_numpy(g["eval/ct_feat"]).float().to(dev)
    assert float((feat - ref).abs().max()) <= OUT_TOL * max(1.0, float(ref.abs().max()))
    assert float((hz.cpu() - torch.from_numpy(g["eval/hazard"]).float()).abs().max()) <= OUT_TOL
    m.train()
    m.rna_encoder[3].p = 0.0; m.fusion[3].p = 0.0
    hz, _ = m(ct, rna, clin, mask)
    assert float((hz.detach().cpu() - torch.from_numpy(g["train/hazard"]).float()).abs().max()) <= OUT_TOL
    (hz * torch.from_numpy(g["train/hazard_weights"]).float().to(dev)).sum().backward()
    errs = {}
    for k, p in m.ct_encoder.named_parameters():
        ref = torch.from_numpy(g["grad/ct_encoder." + k]).to(dev)
        if k in ("0.bias", "3.bias", "6.bias"):              # cancelled by BatchNorm: ~1e-17 in the fp64 reference
            assert float(p.grad.abs().max()) == 0.0 and float(ref.abs().max()) < 1e-9
        else:
            errs[k] = _rel(p.grad, ref)
    # fp64 reference vs bf16 GEMM operands through two bf16 stages and the bf16 head: see GRAD_TOL_FP32 above
    assert max(errs.values()) <= GRAD_TOL_FP32, errs
    for k, v in m.ct_encoder.state_dict().items():
        if "running" in k:
            assert _rel(v, torch.from_numpy(g["sd1/ct_encoder." + k]).to(dev)) <= 5e-3, k
        if "num_batches" in k:
            assert int(v) == int(g["sd1/ct_encoder." + k]) == 1


@pytest.mark.parametrize("cout", [32, 64])
def test_first_conv_primitives(cout):
    """b200surv_ct_conv_first_fwd / _wgrad called directly (the width-32 and the generic weight-gradient kernels)
    against F.conv3d and its weight gradient."""
    import torch.nn.functional as F
    from multimodal_survival_prediction_b200 import _lib as L
    dev = torch.device("cuda", 0)
    L.require_device(0)
    lib = L.load()
    torch.manual_seed(cout)
    B, D, H, W = 3, 11, 16, 9
    x = torch.rand(B, D, H, W, device=dev)
    w = torch.randn(cout, 1, 3, 3, 3, device=dev) * 0.2
    bias = torch.randn(cout, device=dev)
    Do, Ho, Wo = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    R = B * Do * Ho * Wo
    h = torch.empty(R, cout, device=dev)
    st = L.stream_ptr(dev)
    L.check(lib.b200surv_ct_conv_first_fwd(L.ptr(x), L.ptr(w), L.ptr(bias), B, D, H, W, cout, L.ptr(h), st), "fwd")
    wref = w.clone().requires_grad_(True)
    ref = F.conv3d(x.view(B, 1, D, H, W), wref, bias, stride=2, padding=1)           # (B, cout, Do, Ho, Wo)
    assert torch.allclose(h.view(B, Do, Ho, Wo, cout).permute(0, 4, 1, 2, 3), ref, rtol=1e-5, atol=1e-5)
    dy = torch.randn(R, cout, device=dev).bfloat16()
    ref.backward(dy.float().view(B, Do, Ho, Wo, cout).permute(0, 4, 1, 2, 3))
    dw = torch.empty(cout, 27, device=dev)
    ws = torch.empty(lib.b200surv_ct_workspace_bytes(), dtype=torch.uint8, device=dev)
    L.check(lib.b200surv_ct_conv_first_wgrad(L.ptr(x), L.ptr(dy), B, D, H, W, cout, L.ptr(dw), L.ptr(ws), ws.numel(), st), "wgrad")
    assert _rel(dw.view_as(wref), wref.grad) <= 1e-5


def test_graphed_model_step_matches_eager():
    """head.GraphedModelStep (CT encoder + head + Cox loss + backward as one CUDA graph) reproduces the eager step bit
    for bit (dropout off), also on a second batch copied into its static buffers."""
    from multimodal_survival_prediction_b200 import head as ghead, neg_partial_log_likelihood, synth
    dev = torch.device("cuda", 0)
    torch.manual_seed(2)
    B = 4
    m = ghead.MultiModalSurvivalNet().to(dev).train()
    m.rna_encoder[3].p = 0.0; m.fusion[3].p = 0.0
    before = copy.deepcopy(m.state_dict())
    event = torch.tensor([1, 0, 1, 1], device=dev).bool()
    time = torch.tensor([5.0, 3.0, 8.0, 1.0], device=dev)

    def loss_fn(hazard, *_):
        return neg_partial_log_likelihood(hazard, event, time, checks=False)

    batches = []
    for seed in (3, 4):
        _, rna, clin, _ = [t.to(dev) for t in synth.modality_batch(B, seed=seed)]
        batches.append((torch.rand(B, 1, 32, 32, 16, device=dev), rna, clin))
    def eager_step(ct, rna, clin):                    # a function: no tensor of the autograd graph outlives the step
        m.load_state_dict(before); m.zero_grad()      # (a live AccumulateGrad node of the default stream breaks capture)
        loss = loss_fn(m(ct, rna, clin))
        loss.backward()
        return (loss.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}, copy.deepcopy(m.state_dict()))

    eager = [eager_step(*b) for b in batches]         # every step from the same initial state
    m.load_state_dict(before); m.zero_grad()
    step = ghead.GraphedModelStep(m, batches[0][0], batches[0][1], batches[0][2], None, loss_fn)
    for (ct, rna, clin), (loss_e, grads_e, sd_e) in zip(batches, eager):
        m.load_state_dict(before)                     # warm-up and capture ran real steps
        loss, _ = step.step(ct, rna, clin)
        assert torch.equal(loss, loss_e)
        for k, p in m.named_parameters():
            assert torch.equal(p.grad, grads_e[k]), k
        for k, v in m.state_dict().items():
            if "running" in k:
                assert torch.equal(v, sd_e[k]), k


@pytest.mark.parametrize("C", [8, 32, 64])
def test_im2col_col2im_primitives(C):
    """b200surv_ct_im2col against F.unfold-style patch extraction (tap-major columns) and b200surv_ct_col2im as its
    exact adjoint: <im2col(a), d> == <a, col2im(d)>, checked element-wise against the autograd gradient."""
    import torch.nn.functional as F
    from multimodal_survival_prediction_b200 import _lib as L
    dev = torch.device("cuda", 0)
    L.require_device(0)
    lib = L.load()
    torch.manual_seed(C)
    B, D, H, W = 2, 7, 10, 5
    Do, Ho, Wo = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    a = torch.randn(B, D, H, W, C, device=dev).bfloat16()
    col = torch.empty(B * Do * Ho * Wo, 27 * C, dtype=torch.bfloat16, device=dev)
    st = L.stream_ptr(dev)
    L.check(lib.b200surv_ct_im2col(L.ptr(a), B, D, H, W, C, L.ptr(col), st), "im2col")
    af = a.float().requires_grad_(True)
    pad = F.pad(af, (0, 0, 1, 1, 1, 1, 1, 1))                                    # pad W, H, D by 1 (channels last)
    taps = [pad[:, kz:kz + 2 * Do:2, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2, :] for kz in range(3) for ky in range(3) for kx in range(3)]
    ref = torch.stack(taps, dim=4).reshape(B * Do * Ho * Wo, 27 * C)           # (b, zo, yo, xo, tap, c)
    assert torch.equal(col.float(), ref.detach())
    d = torch.randn_like(ref).bfloat16()
    ref.backward(d.float())
    da = torch.empty(B * D * H * W, C, device=dev)
    L.check(lib.b200surv_ct_col2im(L.ptr(d), B, D, H, W, C, L.ptr(da), st), "col2im")
    assert torch.allclose(da.view_as(af), af.grad, rtol=1e-5, atol=1e-5)
