"""the reference's lifelines fallback ``concordance_index(time, -hazard, event)``
(scripts/training/partial_modality_training.py:313-319, scripts/analysis/evaluate_model.py:41-45).

CPU: the restatement of lifelines' sweep (oracle/lifelines_cindex.py) equals the six-counter derivation the product
uses (counters with tolerance 0 on estimate = -score; (C + T/2)/(C + D + T)).  GPU: the product function and the
``lifelines.utils`` shim against the restatement, float64, exact."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cindex as oci
from oracle import lifelines_cindex as oll

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cohort(n, seed, tmax=9, score_ties=True):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, tmax + 1, n).astype(np.float32)
    ev = rng.random(n) < 0.5
    hz = rng.normal(size=n).astype(np.float32)
    if score_ties:
        hz = (np.round(hz * 4) / 4).astype(np.float32)
    return hz, ev, t


@pytest.mark.parametrize("n,seed,tmax", [(2, 0, 1), (5, 1, 2), (40, 2, 3), (300, 3, 9), (300, 4, 5000), (1200, 5, 30)])
def test_lifelines_sweep_equals_six_counter_derivation(n, seed, tmax):
    hz, ev, t = cohort(n, seed, tmax)
    correct, tied, pairs = oll.summary_statistics(t, -hz, ev)
    c = [int(x) for x in oci.counts_brute(hz, ev, t, 0.0)]
    assert correct == c[0] + c[3] and tied == c[2] + c[5] and pairs == sum(c)
    if pairs:
        assert oll.concordance_index(t, -hz, ev) == (correct + tied / 2) / pairs


def test_lifelines_known_answers():
    # KA1: 4 strict comparable pairs, 3 concordant, no same-time pairs -> 0.75
    hz = np.array([0.1, 0.5, -0.3, 0.2]); ev = np.array([1, 0, 1, 1]); t = np.array([5., 3., 8., 1.])
    assert oll.concordance_index(t, -hz, ev) == 0.75
    # all deaths at one time: nothing comparable
    with pytest.raises(ZeroDivisionError):
        oll.concordance_index(np.ones(5), np.arange(5.0), np.ones(5))
    # a censored row at the time of a death IS comparable with it (and only that way round)
    assert oll.summary_statistics([2., 2.], [0.0, 1.0], [1, 0]) == (1, 0, 1)
    assert oll.summary_statistics([2., 2.], [1.0, 0.0], [1, 0]) == (0, 0, 1)
    assert oll.summary_statistics([2., 2.], [1.0, 1.0], [1, 0]) == (0, 1, 1)
    # event_observed=None means everybody died
    assert oll.concordance_index([1., 2., 3.], [1., 2., 3.]) == 1.0
    with pytest.raises(ValueError):
        oll.concordance_index([1., np.nan], [1., 2.], [1, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,tmax", [(2, 10, 1), (37, 11, 3), (348, 12, 9), (5000, 13, 4000), (30_000, 14, 50)])
def test_gpu_lifelines_convention_matches_restatement(n, seed, tmax):
    from multimodal_survival_prediction_b200.cindex import ConcordanceIndex, concordance_index_lifelines
    hz, ev, t = cohort(n, seed, tmax)
    ref = oll.concordance_index(t, -hz, ev)
    # the reference's call: numpy arrays from CPU tensors (partial_modality_training.py:317)
    val = concordance_index_lifelines(torch.from_numpy(t).numpy(), -torch.from_numpy(hz).numpy(), torch.from_numpy(ev).float().numpy())
    assert isinstance(val, float) and val == ref
    obj = ConcordanceIndex(convention="lifelines")
    out = obj(torch.from_numpy(hz), torch.from_numpy(ev), torch.from_numpy(t))
    assert out.dtype == torch.float64 and out.item() == ref
    correct, tied, pairs = oll.summary_statistics(t, -hz, ev)
    c = obj.counts
    assert (c[0] + c[3], c[2] + c[5], sum(c)) == (correct, tied, pairs)


@pytest.mark.gpu
def test_gpu_lifelines_float64_scores_pandas_and_errors():
    """evaluate_model.py:41-45 passes pandas float64 columns: values that are not fp32-representable keep their order and ties."""
    import pandas as pd
    from multimodal_survival_prediction_b200.cindex import concordance_index_lifelines
    rng = np.random.default_rng(5)
    n = 400
    t = rng.integers(1, 30, n).astype(np.float64)
    ev = rng.random(n) < 0.6
    base = np.round(rng.normal(size=n), 1)
    score = base + rng.integers(0, 2, n) * 1e-12           # pairs that differ by 1e-12: distinct in float64, equal in fp32
    df = pd.DataFrame({"survival_time": t, "risk_score": -score, "event": ev.astype(int)})
    val = concordance_index_lifelines(df["survival_time"], -df["risk_score"], df["event"])
    assert val == oll.concordance_index(t, score, ev)
    assert val != oll.concordance_index(t, score.astype(np.float32), ev)       # rounding to fp32 would have changed it
    assert concordance_index_lifelines([1., 2., 3.], [1., 2., 3.]) == 1.0      # event_observed=None
    with pytest.raises(ZeroDivisionError):
        concordance_index_lifelines(np.ones(5), np.arange(5.0), np.ones(5))
    with pytest.raises(ValueError):
        concordance_index_lifelines([1., np.nan], [1., 2.], [1, 1])
    with pytest.raises(ValueError):
        concordance_index_lifelines([1., 2., 3.], [1., 2.], [1, 1, 1])


@pytest.mark.gpu
def test_gpu_lifelines_shim_import_path():
    """``from lifelines.utils import concordance_index`` resolves to the B200 function with PYTHONPATH=shim (the reference's
    import at partial_modality_training.py:316)."""
    code = ("import numpy as np\nfrom lifelines.utils import concordance_index\n"
            "print(concordance_index(np.array([5.,3.,8.,1.],dtype=np.float32), -np.array([0.1,0.5,-0.3,0.2],dtype=np.float32), np.array([1.,0.,1.,1.],dtype=np.float32)))")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shim"), ROOT]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    assert float(r.stdout.strip().splitlines()[-1]) == 0.75
