"""Dormant pin tests: they SKIP in this image (torchsurv / lifelines are absent and cannot be installed: no index, no
wheel -- SURVEY.md 0 fact 1) and light up on the first box that has the real packages.

What they pin, at the call sites the reference uses (scripts/training/partial_modality_training.py:285-294,313-319;
simple_fusion.py:270,311,330-331):

* ``torchsurv.loss.cox.neg_partial_log_likelihood(log_hz, event, time)`` (defaults; and ties_method / reduction
  spelled out) against ``oracle/cox.py`` -- with BOTH readings of Efron's "mean" (`efron_mean_over`) evaluated, the
  recollected default ("event_times") asserted and the other one reported in the failure message;
* ``torchsurv.metrics.cindex.ConcordanceIndex()(estimate, event, time)`` against ``oracle/cindex.py`` ("harrell");
* ``lifelines.utils.concordance_index(time, -hazard, event)`` against ``oracle/lifelines_cindex.py``.

Inputs: the known-answer set KA1-KA5 of SURVEY.md 8c plus a seeded heavy-tie cohort shaped like the headline workload.
The GPU-marked variants compare the product itself (the torchsurv shim's targets) with the real packages.
"""
import os
import sys

import numpy as np
import pytest
import torch

from multimodal_survival_prediction_b200 import synth
from oracle import cindex as oci
from oracle import cox as ocox
from oracle import lifelines_cindex as oll

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim")


def _real(name):
    """Import the REAL package: the repo's shim directory must not shadow it."""
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != SHIM]
    for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
        if "b200surv" in str(getattr(sys.modules[k], "__version__", "")) or SHIM in str(getattr(sys.modules[k], "__file__", "")):
            del sys.modules[k]
    try:
        mod = pytest.importorskip(name, reason=f"{name} is not installed in this image (parity stays unpinned)")
        if "b200surv" in str(getattr(mod, "__version__", "")):
            pytest.skip(f"only the repo's {name} shim is importable")
        return mod
    finally:
        sys.path[:] = saved


def ka_cases():
    rng = np.random.default_rng(7)
    cases = {
        "KA1": (np.array([0.1, 0.5, -0.3, 0.2], np.float32), np.array([1, 0, 1, 1], bool), np.array([5., 3., 8., 1.], np.float32)),
        "KA2_all_tied_all_events": (rng.normal(size=7).astype(np.float32), np.ones(7, bool), np.full(7, 4.0, np.float32)),
        "KA3_constant_eta": (np.zeros(12, np.float32), rng.random(12) < 0.6, np.arange(12, dtype=np.float32)),
        "KA4_reversed_perfect": (-np.arange(16, dtype=np.float32), np.ones(16, bool), np.arange(16, dtype=np.float32)),
        "KA5_small_ties": (rng.normal(size=60).astype(np.float32), rng.random(60) < 0.5, rng.integers(1, 9, 60).astype(np.float32)),
        "KA5_distinct_times": (rng.normal(size=40).astype(np.float32), rng.random(40) < 0.5, rng.permutation(40).astype(np.float32) + 1),
    }
    lh, ev, t = synth.cohort(3000, 1234)                  # heavy ties: integer days, ~30 % events
    cases["heavy_ties_3000"] = (lh.numpy(), ev.numpy(), t.numpy())
    lh, ev, t = synth.cohort(2000, 99, risk_tie_frac=0.2)
    cases["risk_ties_2000"] = (lh.numpy(), ev.numpy(), torch.clamp(torch.floor(t / 100), 1, 40).numpy())
    for k, (a, b, c) in cases.items():
        if not b.any():
            b[0] = True
    return cases


@pytest.mark.parametrize("ties", ["efron", "breslow"])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_torchsurv_cox_pins_the_oracle(ties, reduction):
    _real("torchsurv")
    from torchsurv.loss.cox import neg_partial_log_likelihood as real_nll
    for name, (eta, ev, t) in ka_cases().items():
        x = torch.tensor(eta, dtype=torch.float64, requires_grad=True)
        with __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            loss = real_nll(x, torch.from_numpy(ev), torch.from_numpy(t), ties_method=ties, reduction=reduction)
        loss.backward()
        ours = {k: ocox.cox_nll(eta.astype(np.float64), ev, t, ties, reduction, efron_mean_over=k)
                for k in ("event_times", "events")}
        msg = (f"{name} {ties}/{reduction}: torchsurv {loss.item()!r}; oracle event_times {ours['event_times'][0]!r}, "
               f"events {ours['events'][0]!r}")
        assert abs(loss.item() - ours["event_times"][0]) <= 1e-9 * max(1.0, abs(loss.item())), msg
        if ties == "efron":     # (torchsurv's Breslow builds its denominator with torch.tensor([...]): gradient not compared)
            np.testing.assert_allclose(x.grad.numpy(), ours["event_times"][1], rtol=0, atol=1e-9, err_msg=msg)


def test_torchsurv_cox_default_arguments_are_efron_mean():
    _real("torchsurv")
    from torchsurv.loss.cox import neg_partial_log_likelihood as real_nll
    eta, ev, t = ka_cases()["heavy_ties_3000"]
    with __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        loss = real_nll(torch.tensor(eta, dtype=torch.float64), torch.from_numpy(ev), torch.from_numpy(t))   # the reference's call
    assert abs(loss.item() - ocox.cox_nll(eta.astype(np.float64), ev, t)[0]) <= 1e-9 * abs(loss.item())


def test_torchsurv_cindex_pins_the_oracle():
    _real("torchsurv")
    from torchsurv.metrics.cindex import ConcordanceIndex as RealCI
    for name, (eta, ev, t) in ka_cases().items():
        ref = oci.cindex_from_counts(oci.counts_brute(eta, ev, t, 1e-8), "harrell")
        val = RealCI()(torch.from_numpy(eta), torch.from_numpy(ev), torch.from_numpy(t))
        assert val.dtype == torch.float32, name
        assert val.item() == pytest.approx(ref, abs=1e-6), (name, val.item(), ref)


def test_lifelines_cindex_pins_the_oracle():
    _real("lifelines")
    from lifelines.utils import concordance_index as real_ci
    for name, (eta, ev, t) in ka_cases().items():
        try:
            ref = oll.concordance_index(t, -eta, ev)
        except ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                real_ci(t, -eta, ev)
            continue
        assert real_ci(t, -eta, ev) == ref, name


@pytest.mark.gpu
def test_product_against_real_torchsurv():
    _real("torchsurv")
    from torchsurv.loss.cox import neg_partial_log_likelihood as real_nll
    from torchsurv.metrics.cindex import ConcordanceIndex as RealCI
    import multimodal_survival_prediction_b200 as pkg
    for name, (eta, ev, t) in ka_cases().items():
        x = torch.tensor(eta, requires_grad=True)
        with __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            real_nll(x, torch.from_numpy(ev), torch.from_numpy(t)).backward()
            ref_loss = real_nll(x.detach(), torch.from_numpy(ev), torch.from_numpy(t)).item()
        y = torch.tensor(eta, device="cuda", requires_grad=True)
        loss = pkg.neg_partial_log_likelihood(y, torch.from_numpy(ev).cuda(), torch.from_numpy(t).cuda())
        loss.backward()
        assert abs(loss.item() - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss)), name      # north_star: 1e-5 relative (fp32)
        scale = max(1e-12, float(x.grad.abs().max()))
        assert float((y.grad.cpu() - x.grad).abs().max()) <= 1e-5 * scale + 1e-7, name
        ours = pkg.ConcordanceIndex()(torch.from_numpy(eta), torch.from_numpy(ev), torch.from_numpy(t))
        real = RealCI()(torch.from_numpy(eta), torch.from_numpy(ev), torch.from_numpy(t))
        assert ours.item() == pytest.approx(real.item(), abs=1e-6), name


@pytest.mark.gpu
def test_product_against_real_lifelines():
    _real("lifelines")
    from lifelines.utils import concordance_index as real_ci
    from multimodal_survival_prediction_b200.cindex import concordance_index_lifelines
    for name, (eta, ev, t) in ka_cases().items():
        try:
            ref = real_ci(t, -eta, ev)
        except ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                concordance_index_lifelines(t, -eta, ev)
            continue
        assert concordance_index_lifelines(t, -eta, ev) == ref, name
