"""Harrell's C-index oracle: pair counts (int64) and the ratio.  TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py (parity with torchsurv UNPINNED) and oracle/cindex_oracle.c for the
pair rule.  Three independent implementations, cross-checked in tests/test_oracle.py:

* ``counts_python``  -- literal double loop, the shape of the reference fallback
  (scripts/training/simple_fusion.py:59-73), tiny n only;
* ``counts_brute``   -- C, O(n^2), OpenMP (oracle/cindex_oracle.c:cindex_counts_brute);
* ``counts_fast``    -- C, O(n log n) Fenwick sweep, used at n = 1M.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COUNTER_NAMES = ("conc", "disc", "tied_risk", "conc_st", "disc_st", "tied_st")


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcindex_oracle.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build()
        lib = ctypes.CDLL(path)
        i64, f32p, u8p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p
        lib.cindex_counts_brute.argtypes = [f32p, f32p, u8p, i64, ctypes.c_float, i64, i64, ctypes.c_void_p]
        lib.cindex_counts_brute.restype = None
        lib.cindex_counts_fast.argtypes = [f32p, f32p, u8p, i64, ctypes.c_float, ctypes.c_void_p]
        lib.cindex_counts_fast.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _prep(est, event, time):
    est = np.ascontiguousarray(est, dtype=np.float32)
    time = np.ascontiguousarray(time, dtype=np.float32)
    event = np.ascontiguousarray(np.asarray(event).astype(bool), dtype=np.uint8)
    assert est.shape == time.shape == event.shape and est.ndim == 1
    return est, event, time


def counts_python(est, event, time, tied_tol=1e-8):
    est, event, time = _prep(est, event, time)
    tol = np.float32(tied_tol)
    out = np.zeros(6, dtype=np.int64)
    n = len(est)
    for i in range(n):
        if not event[i]:
            continue
        for j in range(n):
            strict = time[j] > time[i]
            same = (time[j] == time[i]) and not event[j]
            if not (strict or same):
                continue
            tie = np.abs(np.float32(est[i] - est[j])) <= tol
            conc = (not tie) and est[j] < est[i]
            k = 2 if tie else (0 if conc else 1)
            out[k + (0 if strict else 3)] += 1
    return out


def counts_brute(est, event, time, tied_tol=1e-8, row_begin=0, row_end=None):
    est, event, time = _prep(est, event, time)
    n = len(est)
    row_end = n if row_end is None else row_end
    out = np.zeros(6, dtype=np.int64)
    _lib().cindex_counts_brute(est.ctypes.data, time.ctypes.data, event.ctypes.data, n,
                               ctypes.c_float(np.float32(tied_tol)), row_begin, row_end, out.ctypes.data)
    return out


def counts_fast(est, event, time, tied_tol=1e-8):
    est, event, time = _prep(est, event, time)
    out = np.zeros(6, dtype=np.int64)
    rc = _lib().cindex_counts_fast(est.ctypes.data, time.ctypes.data, event.ctypes.data, len(est),
                                   ctypes.c_float(np.float32(tied_tol)), out.ctypes.data)
    if rc != 0:
        raise MemoryError("cindex_counts_fast")
    return out


def cindex_from_counts(counts, convention="harrell"):
    """Ratio in float64.  'harrell' = (C + T/2)/(C + D + T) over strict + same-time pairs
    (scikit-survival / torchsurv-style, recollected); 'fallback' = the reference's in-repo
    rule, simple_fusion.py:59-73: strict pairs only, no credit for risk ties, 0.5 if none."""
    c = [int(x) for x in counts]
    if convention == "harrell":
        C, D, T = c[0] + c[3], c[1] + c[4], c[2] + c[5]
        den = C + D + T
        return (C + 0.5 * T) / den if den > 0 else 0.5
    if convention == "fallback":
        den = c[0] + c[1] + c[2]
        return c[0] / den if den > 0 else 0.5
    raise ValueError(convention)
