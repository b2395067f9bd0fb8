"""Definition-level Cox partial log-likelihood oracle (pure Python / numpy loops, O(n^2)).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Use only for small n.

This file writes the textbook formulas down literally, one risk set at a time, so the
vectorised oracle (oracle/cox.py) and the CUDA kernels have something obviously correct
to be compared with.  Structure follows the call site contract of the reference:

* reference call sites (3 positional args, defaults decide ties/reduction):
  scripts/training/partial_modality_training.py:285-288, simple_fusion.py:270,311,
  final_multimodal.py:158-162, flexible_multimodal.py:290,332, train_rnaseq_only.py:169,193
* arithmetic lives in torchsurv (requirements.txt:32), absent here => PARITY UNPINNED.
  Published algorithm restated:
    no ties : l = sum_{i: d_i=1} [eta_i - log sum_{j: t_j >= t_i} exp(eta_j)]
    Breslow : same formula, tied rows share the full risk set {t_j >= t_i}
    Efron   : per distinct event time t with tied events H (|H| = m), risk set R:
              sum_{i in H} eta_i - sum_{l=0}^{m-1} log( sum_R exp(eta) - (l/m) sum_H exp(eta) )
  Conventions recollected from torchsurv, each an explicit option here:
    - default ties_method="efron", reduction="mean";
    - "mean" averages the vector of terms the method produces: one per EVENT for
      no-ties/Breslow, one per DISTINCT EVENT TIME for Efron (efron_mean_over);
    - zero events (or empty input) -> loss 0.0.
"""
from __future__ import annotations

import math

import numpy as np


def cox_terms_def(log_hz, event, time, ties_method="efron"):
    """Return the list of partial-log-likelihood terms (float64), textbook definition."""
    eta = np.asarray(log_hz, dtype=np.float64)
    ev = np.asarray(event).astype(bool)
    t = np.asarray(time)
    n = eta.shape[0]
    terms = []
    if ties_method == "breslow":
        for i in range(n):
            if not ev[i]:
                continue
            denom = 0.0
            for j in range(n):
                if t[j] >= t[i]:
                    denom += math.exp(eta[j])
            terms.append(eta[i] - math.log(denom))
    elif ties_method == "efron":
        for tu in sorted(set(t.tolist())):
            H = [i for i in range(n) if t[i] == tu and ev[i]]
            if not H:
                continue
            R = [j for j in range(n) if t[j] >= tu]
            m = len(H)
            d_all = sum(math.exp(eta[j]) for j in R)
            d_tie = sum(math.exp(eta[i]) for i in H)
            val = sum(eta[i] for i in H)
            for l in range(m):
                val -= math.log(d_all - (l / m) * d_tie)
            terms.append(val)
    else:
        raise ValueError(f"ties_method {ties_method!r}")
    return terms


def cox_nll_def(log_hz, event, time, ties_method="efron", reduction="mean",
                efron_mean_over="event_times"):
    """Negative partial log-likelihood, definition-level.  Returns a Python float."""
    ev = np.asarray(event).astype(bool)
    if len(ev) == 0 or ev.sum() == 0:
        return 0.0
    terms = cox_terms_def(log_hz, event, time, ties_method)
    total = -float(np.sum(terms))
    if reduction == "sum":
        return total
    if reduction != "mean":
        raise ValueError(f"reduction {reduction!r}")
    if ties_method == "efron" and efron_mean_over == "event_times":
        return total / len(terms)
    return total / int(ev.sum())


def cox_grad_fd(log_hz, event, time, h=1e-6, **kw):
    """Central finite-difference gradient of cox_nll_def (float64) -- checks the analytic one."""
    eta = np.asarray(log_hz, dtype=np.float64).copy()
    g = np.zeros_like(eta)
    for i in range(eta.shape[0]):
        e0 = eta[i]
        eta[i] = e0 + h
        fp = cox_nll_def(eta, event, time, **kw)
        eta[i] = e0 - h
        fm = cox_nll_def(eta, event, time, **kw)
        eta[i] = e0
        g[i] = (fp - fm) / (2 * h)
    return g
