"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the Python host layer validates arguments and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200surv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200surv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200surv.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(L.SIGNATURES) <= set(names)


def test_version_and_sizes_without_gpu():
    lib = L.load()
    assert lib.b200surv_version() >= 1
    assert ctypes.sizeof(L.CoxHeader) == 64
    assert lib.b200surv_cox_state_bytes(1000, 1, L.COX_BINNED, 4096) == 64 + 8 * 4096
    assert lib.b200surv_cox_state_bytes(1000, 3, L.COX_SMALL, 0) == 3 * 64 + 4 * 1000
    assert lib.b200surv_cox_bins_sum_count(4096) == 3 * 4096 + 4
    assert lib.b200surv_cox_workspace_bytes(1 << 24, 1, L.COX_BINNED, 4096) > 148 * 20 * 4096


def test_bad_arguments_are_rejected_on_the_host():
    lib = L.load()
    # null pointers / bad sizes are rejected before any CUDA call
    rc = lib.b200surv_cox_fwd(None, None, None, None, 10, 1, 2, 0, L.COX_BINNED, 4096, 0.0, None, None, 0, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.b200surv_last_error()
    rc = lib.b200surv_cindex_counts(None, None, None, None, 10, 1, 0, 10, 1e-8, 1, None, None, 0, None)
    assert rc == -1


def test_python_validation_mirrors_torchsurv_contract():
    x = torch.randn(5); t = torch.rand(5); e = torch.tensor([1, 0, 1, 1, 0])
    with pytest.raises(ValueError):      # reference README.md:318-321: events must be boolean
        pkg.neg_partial_log_likelihood(x, e, t)
    with pytest.raises(ValueError):
        pkg.neg_partial_log_likelihood(x, e.bool(), t[:4])
    with pytest.raises(ValueError):
        pkg.neg_partial_log_likelihood(x, e.bool(), t, ties_method="exact")
    with pytest.raises(ValueError):
        pkg.neg_partial_log_likelihood(x, e.bool(), t, reduction="median")
    with pytest.raises(ValueError):
        pkg.ConcordanceIndex()(x, e, t)
    with pytest.raises(NotImplementedError):
        pkg.ConcordanceIndex()(x, e.bool(), t, weight=torch.ones(5))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    x = torch.randn(5, requires_grad=True); t = torch.rand(5); e = torch.tensor([1, 0, 1, 1, 0]).bool()
    with pytest.raises(pkg.B200SurvError):
        pkg.neg_partial_log_likelihood(x, e, t)
    with pytest.raises(pkg.B200SurvError):
        pkg.ConcordanceIndex()(x, e, t)


def test_shim_resolves_to_b200_kernels():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    try:
        cox = importlib.import_module("torchsurv.loss.cox")
        ci = importlib.import_module("torchsurv.metrics.cindex")
        assert cox.neg_partial_log_likelihood is pkg.neg_partial_log_likelihood
        assert ci.ConcordanceIndex is pkg.ConcordanceIndex
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))
        for m in [m for m in sys.modules if m.startswith("torchsurv")]:
            del sys.modules[m]


def test_ct_encoder_and_validation_cohort_have_no_cpu_path():
    """The 8f widenings fail loudly without a B200, like the rest of the package; sizes are computable on the host."""
    import torch
    from multimodal_survival_prediction_b200 import B200SurvError, ValidationCohort, _lib as L
    from multimodal_survival_prediction_b200.ctenc import CTEncoderCNN
    lib = L.load()
    assert lib.b200surv_ct_encoder_saved_bytes(4, 64, 64, 32) > 4 * 16384 * 32 * 4      # at least the stage-1 activations
    assert lib.b200surv_ct_encoder_workspace_bytes(4, 64, 64, 32) > 0
    assert lib.b200surv_ct_encoder_saved_bytes(0, 64, 64, 32) == 0
    assert 1 <= lib.b200surv_gemm_splitk_slices(64, 864, 32768) <= 32
    enc = CTEncoderCNN()
    assert [k for k in enc.state_dict()][:2] == ["0.weight", "0.bias"] and enc[0].weight.shape == (32, 1, 3, 3, 3)
    with pytest.raises(B200SurvError):
        enc(torch.zeros(2, 1, 8, 8, 8))
    with pytest.raises(B200SurvError):
        ValidationCohort(capacity=8, device="cpu")
