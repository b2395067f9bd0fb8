"""lifelines.utils.concordance_index restated on the CPU.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference's second C-index fallback is ``concordance_index(time, -hazard, event)`` from lifelines
(scripts/training/partial_modality_training.py:313-319; scripts/analysis/evaluate_model.py:41-45; dependency
``lifelines>=0.27.0``, requirements.txt:33, a floor without an exact pin).  lifelines is absent from this image and
cannot be installed, and the reference holds no test at that call site, so this file restates the algorithm lifelines
publishes (``lifelines/utils/concordance.py``: ``concordance_index`` -> ``_concordance_summary_statistics`` ->
``_handle_pairs``) from its documentation and my reading of it; ``tests/test_torchsurv_pin.py`` compares it with the
real package whenever a box has it.

The published algorithm: rows are visited in order of exit time; all rows that DIED at a time are compared with the
pool of predictions of the rows that died strictly earlier and are then added to the pool; the rows CENSORED at that
time are handled after them (so they also meet the deaths of their own time) and never enter the pool.  A pair (pool
row i, visited row j) is correct when ``pred_i < pred_j``, tied when ``pred_i == pred_j`` (exact, float64).
``concordance_index = (correct + tied / 2) / pairs``; zero pairs raises ZeroDivisionError.
"""
from __future__ import annotations

import bisect

import numpy as np


def summary_statistics(event_times, predicted_scores, event_observed=None):
    """(num_correct, num_tied, num_pairs) by the sweep described above (a sorted list stands in for lifelines' tree)."""
    t = np.asarray(event_times, dtype=float).reshape(-1)
    p = np.asarray(predicted_scores, dtype=float).reshape(-1)
    e = np.ones(t.shape[0], dtype=bool) if event_observed is None else np.asarray(event_observed).astype(float).reshape(-1) != 0
    if not (t.shape == p.shape == e.shape):
        raise ValueError("Observed events must be 1-dimensional of same length as event times")
    if np.isnan(t).any() or np.isnan(p).any():
        raise ValueError("NaNs detected in inputs, please correct or drop.")
    order_d = np.argsort(t[e], kind="stable")
    died_t, died_p = t[e][order_d], p[e][order_d]
    order_c = np.argsort(t[~e], kind="stable")
    cens_t, cens_p = t[~e][order_c], p[~e][order_c]
    pool: list = []
    correct = tied = pairs = 0
    di = ci = 0

    def handle(truth, pred, first):
        nxt = first
        while nxt < len(truth) and truth[nxt] == truth[first]:
            nxt += 1
        c = k = 0
        for i in range(first, nxt):
            lo = bisect.bisect_left(pool, pred[i])
            hi = bisect.bisect_right(pool, pred[i])
            c += lo
            k += hi - lo
        return len(pool) * (nxt - first), c, k, nxt

    while di < len(died_t) or ci < len(cens_t):
        more_c, more_d = ci < len(cens_t), di < len(died_t)
        if more_c and (not more_d or died_t[di] > cens_t[ci]):
            n_p, c, k, ci = handle(cens_t, cens_p, ci)
        else:
            n_p, c, k, nxt = handle(died_t, died_p, di)
            for v in died_p[di:nxt]:
                bisect.insort(pool, v)
            di = nxt
        pairs += n_p
        correct += c
        tied += k
    return correct, tied, pairs


def concordance_index(event_times, predicted_scores, event_observed=None) -> float:
    correct, tied, pairs = summary_statistics(event_times, predicted_scores, event_observed)
    if pairs == 0:
        raise ZeroDivisionError("No admissable pairs in the dataset.")
    return (correct + tied / 2) / pairs
