"""CPU tests of the oracle itself: pinned against the golden vectors produced by the reference's
runnable fallback code (oracle/gen_golden.py) and against definition-level restatements."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import cindex as oci
from oracle import cox as ocox
from oracle import cox_def
from oracle import head as ohead


# ---------------------------------------------------------------- Cox
def test_cox_oracle_matches_reference_fallback_on_tie_free_inputs(golden):
    g = golden("cox_fallback.npz")
    for c in g["cases"]:
        eta, ev, t = g[f"{c}/log_hz"], g[f"{c}/event"], g[f"{c}/time"]
        assert len(np.unique(t)) == len(t)
        for ties in ("efron", "breslow"):
            # the fallback divides by the event count -> efron_mean_over="events" (identical without ties)
            loss, grad = ocox.cox_nll(eta, ev, t, ties_method=ties, efron_mean_over="events")
            for variant in ("partial_modality", "simple_fusion", "rnaseq_only"):
                ref_l = float(g[f"{c}/{variant}/f64/loss"])
                ref_g = g[f"{c}/{variant}/f64/grad"]
                assert abs(loss - ref_l) <= 1e-7 * max(1.0, abs(ref_l)), (c, ties, variant)
                np.testing.assert_allclose(grad, ref_g, rtol=0, atol=1e-7)


def test_cox_ka1_known_answer():
    # SURVEY.md 8c KA1 (hand-verified): no-ties mean-over-events NLL and gradient
    eta = np.array([0.1, 0.5, -0.3, 0.2]); ev = np.array([1, 0, 1, 1], bool); t = np.array([5., 3., 8., 1.])
    loss, grad = ocox.cox_nll(eta, ev, t)
    assert abs(loss - 0.6213334104377958) < 1e-8
    np.testing.assert_allclose(grad, [-0.0556576616101615, 0.11653107916781796, 0.1861315690524536,
                                      -0.24700498661011], atol=1e-8)


def test_cox_ka2_all_tied_all_events():
    rng = np.random.default_rng(3)
    eta = rng.normal(size=7); ev = np.ones(7, bool); t = np.full(7, 4.0)
    S = np.exp(eta).sum(); m = 7
    br, _ = ocox.cox_nll(eta, ev, t, "breslow", "sum")
    ef, _ = ocox.cox_nll(eta, ev, t, "efron", "sum")
    assert abs(br - -(eta.sum() - m * np.log(S))) < 1e-10
    assert abs(ef - -(eta.sum() - sum(np.log((1 - l / m) * S) for l in range(m)))) < 1e-10
    # Efron yields ONE term here: "mean" over event times leaves the sum unchanged
    ef_mean, _ = ocox.cox_nll(eta, ev, t, "efron", "mean")
    assert abs(ef_mean - ef) < 1e-12
    ef_mean_ev, _ = ocox.cox_nll(eta, ev, t, "efron", "mean", efron_mean_over="events")
    assert abs(ef_mean_ev - ef / m) < 1e-12


@pytest.mark.parametrize("ties", ["efron", "breslow"])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_cox_vectorised_equals_definition_and_fd_gradient(ties, reduction):
    rng = np.random.default_rng(0)
    for n, tmax in ((1, 3), (2, 2), (17, 4), (60, 8), (60, 1000)):
        t = rng.integers(1, tmax + 1, n).astype(np.float32)
        ev = rng.random(n) < 0.5
        eta = rng.normal(size=n)
        loss, grad = ocox.cox_nll(eta, ev, t, ties, reduction)
        ref = cox_def.cox_nll_def(eta, ev, t, ties, reduction)
        assert abs(loss - ref) <= 1e-10 * max(1.0, abs(ref))
        if ev.any():
            fd = cox_def.cox_grad_fd(eta, ev, t, ties_method=ties, reduction=reduction)
            np.testing.assert_allclose(grad, fd, atol=5e-7 * max(1.0, np.abs(fd).max()))


def test_cox_no_events_and_empty():
    assert ocox.cox_nll(np.zeros(3), np.zeros(3, bool), np.arange(3.0))[0] == 0.0
    assert ocox.cox_nll(np.zeros(0), np.zeros(0, bool), np.zeros(0))[0] == 0.0


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2 ** 31 - 1), st.floats(-3, 3))
def test_cox_properties(n, seed, shift):
    rng = np.random.default_rng(seed)
    t = rng.integers(1, 6, n).astype(np.float32); ev = rng.random(n) < 0.6; eta = rng.normal(size=n)
    if not ev.any():
        ev[0] = True
    for ties in ("efron", "breslow"):
        l0, g0 = ocox.cox_nll(eta, ev, t, ties, "sum")
        l1, _ = ocox.cox_nll(eta + shift, ev, t, ties, "sum")          # invariance to eta + c
        assert abs(l0 - l1) < 1e-8 * max(1, abs(l0))
        assert abs(g0.sum()) < 1e-9 * max(1, np.abs(g0).sum())         # sum of gradient = 0
        p = rng.permutation(n)                                         # permutation invariance
        l2, g2 = ocox.cox_nll(eta[p], ev[p], t[p], ties, "sum")
        assert abs(l0 - l2) < 1e-9 * max(1, abs(l0))
        np.testing.assert_allclose(g0[p], g2, atol=1e-10)
    td = rng.permutation(1000)[:n].astype(np.float32)                  # distinct times: methods agree
    a, _ = ocox.cox_nll(eta, ev, td, "efron", "mean")
    b, _ = ocox.cox_nll(eta, ev, td, "breslow", "mean")
    assert abs(a - b) < 1e-10 * max(1, abs(a))


def test_cox_segmented():
    rng = np.random.default_rng(5)
    n = 50; t = rng.integers(1, 9, n).astype(np.float32); ev = rng.random(n) < 0.5; eta = rng.normal(size=n)
    off = [0, 11, 30, 50]
    ls, g = ocox.cox_nll_segmented(eta, ev, t, off)
    for s in range(3):
        l, gg = ocox.cox_nll(eta[off[s]:off[s + 1]], ev[off[s]:off[s + 1]], t[off[s]:off[s + 1]])
        assert l == ls[s]
        np.testing.assert_array_equal(gg, g[off[s]:off[s + 1]])


# ---------------------------------------------------------------- C-index
def test_cindex_oracle_matches_reference_fallback(golden):
    g = golden("cindex_fallback.npz")
    for c in g["cases"]:
        est, ev, t = g[f"{c}/est"], g[f"{c}/event"], g[f"{c}/time"]
        # the fallback compares hazards with a strict '>' -> tied_tol = 0
        a = oci.counts_python(est, ev, t, 0.0)
        b = oci.counts_brute(est, ev, t, 0.0)
        f = oci.counts_fast(est, ev, t, 0.0)
        assert (a == b).all() and (a == f).all(), c
        val = np.float32(oci.cindex_from_counts(a, "fallback"))
        assert val == g[f"{c}/value"], c


def test_cindex_ka_cases():
    est = np.array([0.1, 0.5, -0.3, 0.2], np.float32); ev = np.array([1, 0, 1, 1], bool)
    t = np.array([5, 3, 8, 1], np.float32)
    c = oci.counts_brute(est, ev, t)
    assert list(c) == [3, 1, 0, 0, 0, 0] and oci.cindex_from_counts(c, "fallback") == 0.75
    # KA3: constant estimate -> every comparable pair is a risk tie
    c = oci.counts_brute(np.zeros(6, np.float32), np.ones(6, bool), np.arange(6, dtype=np.float32))
    assert c[0] == c[1] == 0 and c[2] == 15
    assert oci.cindex_from_counts(c, "harrell") == 0.5 and oci.cindex_from_counts(c, "fallback") == 0.0
    # KA4: perfectly ranked (higher estimate = earlier event)
    c = oci.counts_brute(-np.arange(6, dtype=np.float32), np.ones(6, bool), np.arange(6, dtype=np.float32))
    assert c[0] == 15 and c[1] == c[2] == 0
    # KA2: all times equal, all events -> nothing comparable
    c = oci.counts_brute(np.arange(5, dtype=np.float32), np.ones(5, bool), np.ones(5, np.float32))
    assert c.sum() == 0 and oci.cindex_from_counts(c) == 0.5


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 120), st.integers(0, 2 ** 31 - 1), st.sampled_from([0.0, 1e-8, 0.3]))
def test_cindex_three_implementations_agree(n, seed, tol):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 7, n).astype(np.float32); ev = rng.random(n) < 0.5
    est = np.round(rng.normal(size=n) * 3).astype(np.float32) / 3 if seed % 2 else rng.normal(size=n).astype(np.float32)
    a = oci.counts_python(est, ev, t, tol); b = oci.counts_brute(est, ev, t, tol); f = oci.counts_fast(est, ev, t, tol)
    assert (a == b).all() and (a == f).all()
    p = rng.permutation(n)
    assert (oci.counts_fast(est[p], ev[p], t[p], tol) == a).all()
    # row sharding adds up
    h = n // 2
    assert (oci.counts_brute(est, ev, t, tol, 0, h) + oci.counts_brute(est, ev, t, tol, h, n) == a).all()


def test_cindex_fast_equals_brute_moderate_n():
    rng = np.random.default_rng(9)
    n = 20000
    t = np.clip(np.floor(rng.exponential(1000, n)), 1, 4000).astype(np.float32)
    ev = rng.random(n) < 0.3
    est = rng.normal(size=n).astype(np.float32)
    sel = rng.random(n) < 0.1
    est[sel] = np.round(est[sel] * 100) / 100
    assert (oci.counts_fast(est, ev, t) == oci.counts_brute(est, ev, t)).all()


# ---------------------------------------------------------------- head
@pytest.mark.parametrize("tag", ["gated", "ungated"])
def test_head_oracle_matches_reference_modules(golden, tag):
    g = golden(f"head_{tag}.npz")
    p = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0/")}
    rna, clin = torch.from_numpy(g["rna"]), torch.from_numpy(g["clinical"])
    mask = torch.from_numpy(g["mask"]) if tag == "gated" else None
    # eval mode
    out = ohead.head_forward(p, torch.from_numpy(g["eval/ct_feat"]), rna, clin, mask, train=False)
    hz = out[0] if tag == "gated" else out
    np.testing.assert_allclose(hz.numpy(), g["eval/hazard"], rtol=1e-12, atol=1e-12)
    if tag == "gated":
        np.testing.assert_allclose(out[1].numpy(), g["eval/gate"], rtol=1e-12, atol=1e-12)
        assert abs(float(ohead.gate_entropy_loss(out[1])) - float(g["eval/gate_entropy"])) < 1e-12
    # train mode (dropout off): outputs, parameter gradients, running statistics
    pt = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in p.items()}
    stats = {}
    out = ohead.head_forward(pt, torch.from_numpy(g["train/ct_feat"]), rna, clin, mask, train=True, stats_out=stats)
    hz = out[0] if tag == "gated" else out
    np.testing.assert_allclose(hz.detach().numpy(), g["train/hazard"], rtol=1e-10, atol=1e-10)
    obj = (hz * torch.from_numpy(g["train/hazard_weights"])).sum()
    if tag == "gated":
        obj = obj + 0.01 * ohead.gate_entropy_loss(out[1])
    obj.backward()
    for k in g.files:
        if k.startswith("grad/"):
            np.testing.assert_allclose(pt[k[5:]].grad.numpy(), g[k], rtol=1e-8, atol=1e-10, err_msg=k)
    pd = {k: v.detach() for k, v in pt.items()}
    ohead.bn_running_update(pd, stats)
    for k in g.files:
        if k.startswith("sd1/") and "running" in k:
            np.testing.assert_allclose(pd[k[4:]].numpy(), g[k], rtol=1e-10, atol=1e-12, err_msg=k)


def test_head_state_dict_manifest(golden):
    # contract a1: key names and shapes of the full-size reference modules
    g = golden("head_gated_manifest.npz")
    assert tuple(g["rna_encoder.0.weight"]) == (512, 5005)
    assert tuple(g["gate.0.weight"]) == (64, 291) and tuple(g["gate.2.weight"]) == (3, 64)
    assert tuple(g["fusion.0.weight"]) == (256, 288) and tuple(g["cox_head.weight"]) == (1, 128)
    p = ohead.init_head_params(5005, gated=True)
    for k, v in p.items():
        assert tuple(g[k]) == tuple(v.shape), k
    u = golden("head_ungated_manifest.npz")
    assert "gate.0.weight" not in u.files and tuple(u["fusion.0.weight"]) == (256, 288)


def test_ct_encoder_restatement_matches_reference_class():
    """oracle/ctenc.py (the CNN the GPU tests compare against) reproduces what the reference's own class computed:
    tests/golden/ct_encoder.npz, eval features and training-step gradients of ct_encoder.* (fp64)."""
    import os
    from oracle.ctenc import reference_cnn
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ct_encoder.npz"))
    cnn = reference_cnn().double()
    cnn.load_state_dict({k[len("sd0/ct_encoder."):]: torch.from_numpy(g[k]).double() if g[k].dtype.kind == "f" else torch.from_numpy(g[k])
                         for k in g.files if k.startswith("sd0/ct_encoder.")})
    ct = torch.from_numpy(g["ct"]).double()
    cnn.eval()
    with torch.no_grad():
        feat = cnn(ct).view(ct.shape[0], -1)
    assert np.abs(feat.numpy() - g["eval/ct_feat"]).max() <= 1e-6      # the fixture stores the parameters as float32
    cnn.train()
    cnn(ct)
    for k, v in cnn.state_dict().items():
        if "running" in k:
            assert np.abs(v.numpy() - g["sd1/ct_encoder." + k]).max() <= 1e-6, k


def test_cpu_model_port_has_the_reference_state_dict():
    """oracle/model_torch.MultiModalNetCPU (bench.py's CPU leg of configs[0]) has exactly the keys and shapes of the
    reference's MultiModalSurvivalNet (manifest written from the reference class by oracle/gen_golden.py), and one
    training step of it runs and changes the parameters."""
    import os
    from oracle import model_torch
    man = np.load(os.path.join(os.path.dirname(__file__), "golden", "head_ungated_manifest.npz"))
    m = model_torch.MultiModalNetCPU()
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(man.files)
    for k in man.files:
        assert tuple(sd[k].shape) == tuple(int(v) for v in man[k]), k
    torch.manual_seed(0)
    small = model_torch.MultiModalNetCPU(rna_dim=16).train()
    opt = torch.optim.AdamW(small.parameters(), lr=1e-3)
    w0 = small.cox_head.weight.detach().clone()
    loss = model_torch.training_step(small, opt, torch.rand(4, 1, 16, 16, 8), torch.randn(4, 16), torch.rand(4, 1),
                                     torch.tensor([1, 0, 1, 1]).bool(), torch.tensor([5.0, 3.0, 8.0, 1.0]))
    assert np.isfinite(loss) and not torch.equal(w0, small.cox_head.weight.detach())
