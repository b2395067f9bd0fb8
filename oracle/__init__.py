"""CPU oracle for the survival hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or as the timed CPU baseline.
The product package (``multimodal_survival_prediction_b200``) never imports it and
fails loudly when its CUDA library is missing.

PARITY STATUS -- read before trusting a number:

* Cox partial likelihood / C-index as defined by ``torchsurv``: **parity unpinned**.
  torchsurv (requirements.txt:32, ``torchsurv>=0.1.0``, no exact pin) is not vendored
  in the reference, not installed in the build image and not installable (no index).
  The reference holds no test or golden vector at that boundary.  ``oracle/cox.py`` and
  ``oracle/cindex.py`` restate the published textbook algorithms (Breslow 1974, Efron
  1977, Harrell 1982 with the scikit-survival comparability rule torchsurv documents)
  and expose each unverifiable convention as an explicit option.
* What *is* pinned: the reference's own runnable fallback loss / C-index and its model
  classes.  ``oracle/gen_golden.py`` AST-extracts those from ``/root/reference`` in the
  build container, runs them on seeded inputs and stores inputs+outputs under
  ``tests/golden/``; the oracle is checked against every one of those vectors
  (tie-free inputs for the loss, where the fallback equals the textbook formula).
"""
