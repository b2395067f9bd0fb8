"""GPU parity tests of the tcgen05/TMEM bf16 GEMM primitive (through the C ABI) vs torch fp32 matmul
on the same bf16-rounded operands.  Tolerance: fp32 accumulation order only -> 1e-4 relative to the
row scale (the operands are identical bf16 values on both sides)."""
import ctypes

import pytest
import torch

from multimodal_survival_prediction_b200 import _lib as L

pytestmark = pytest.mark.gpu


def gemm(a, a_mn, b, b_mn, M, N, K, bias=None, relu=False, want_bf16=False):
    lib = L.load()
    dev = a.device
    L.require_device(dev.index)
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    cb = torch.zeros((M, (N + 7) // 8 * 8), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    rc = lib.b200surv_gemm_bf16(L.ptr(a), a.stride(0), int(a_mn), L.ptr(b), b.stride(0), int(b_mn), M, N, K, L.ptr(c),
                                c.stride(0), L.ptr(cb), cb.stride(0) if cb is not None else 0, L.ptr(bias), int(relu),
                                L.stream_ptr(dev))
    L.check(rc, "b200surv_gemm_bf16")
    return c, cb


def padded(rows, cols, seed):
    """bf16 matrix [rows][cols] stored with a pitch that is a multiple of 8 elements (TMA requirement)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ld = (cols + 7) // 8 * 8
    buf = torch.zeros(rows, ld, dtype=torch.bfloat16)
    buf[:, :cols] = torch.randn(rows, cols, generator=g).to(torch.bfloat16)
    return buf.cuda()[:, :cols]


SHAPES = [(128, 128, 64), (128, 128, 512), (256, 384, 320), (4, 512, 5005), (4096, 512, 5005), (100, 37, 291),
          (512, 5005, 4096), (300, 128, 512), (129, 130, 65), (1, 8, 8)]


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
def test_gemm_layouts_and_ragged_shapes(a_mn, b_mn):
    for i, (M, N, K) in enumerate(SHAPES):
        a = padded(K, M, 10 + i) if a_mn else padded(M, K, 10 + i)
        b = padded(K, N, 50 + i) if b_mn else padded(N, K, 50 + i)
        A = a.float().t() if a_mn else a.float()          # [M][K]
        Bm = b.float() if b_mn else b.float().t()         # [K][N]
        ref = A @ Bm
        c, _ = gemm(a, a_mn, b, b_mn, M, N, K)
        scale = ref.abs().max().item() + 1e-6
        err = (c - ref).abs().max().item()
        assert err <= 2e-4 * scale, (M, N, K, a_mn, b_mn, err, scale)


def test_gemm_epilogue_bias_relu_bf16():
    M, N, K = 300, 200, 136
    a, b = padded(M, K, 1), padded(N, K, 2)
    bias = torch.randn(N, device="cuda")
    ref = torch.relu(a.float() @ b.float().t() + bias)
    c, cb = gemm(a, False, b, False, M, N, K, bias=bias, relu=True, want_bf16=True)
    assert (c - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert (cb[:, :N].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert torch.all(cb[:, N:] == 0)


@pytest.mark.parametrize("tile_n", [192, 256, 512])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_wide_tiles(tile_n, a_mn, b_mn):
    """b200surv_gemm_bf16_ex: 128 x 192 and 128 x 256 tiles (TMEM accumulators of 256 columns) and, tile_n = 512, CTA pairs on
    256 x 256 tiles (tcgen05.mma.cta_group::2); ragged shapes, epilogue."""
    lib = L.load()
    for i, (M, N, K) in enumerate([(128, 256, 64), (4096, 512, 5005), (512, 5005, 4096), (300, 200, 136), (129, 385, 65), (4, 512, 5005)]):
        a = padded(K, M, 20 + i) if a_mn else padded(M, K, 20 + i)
        b = padded(K, N, 70 + i) if b_mn else padded(N, K, 70 + i)
        A = a.float().t() if a_mn else a.float()
        Bm = b.float() if b_mn else b.float().t()
        bias = torch.randn(N, device="cuda")
        ref = torch.relu(A @ Bm + bias)
        c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
        cb = torch.zeros((M, (N + 7) // 8 * 8), dtype=torch.bfloat16, device="cuda")
        for want_bf16 in (False, True):       # one output: staged, coalesced epilogue; two outputs: direct stores
            rc = lib.b200surv_gemm_bf16_ex(L.ptr(a), a.stride(0), int(a_mn), L.ptr(b), b.stride(0), int(b_mn), M, N, K, L.ptr(c),
                                           c.stride(0), L.ptr(cb) if want_bf16 else None, cb.stride(0) if want_bf16 else 0,
                                           L.ptr(bias), 1, tile_n, 1, None, L.stream_ptr(a.device))
            L.check(rc, "b200surv_gemm_bf16_ex")
            scale = ref.abs().max().item() + 1e-6
            assert (c - ref).abs().max().item() <= 2e-4 * scale, (M, N, K, tile_n, a_mn, b_mn, want_bf16)
            if want_bf16:
                assert (cb[:, :N].float() - ref).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize("tile_n,splits", [(256, 2), (128, 3), (192, 4), (512, 2)])
def test_gemm_forced_k_split(tile_n, splits):
    lib = L.load()
    M, N, K = 4096, 512, 5005
    a, b = padded(M, K, 5), padded(N, K, 6)
    ref = a.float() @ b.float().t()
    sl = torch.full((splits, M, N), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.b200surv_gemm_bf16_ex(L.ptr(a), a.stride(0), 0, L.ptr(b), b.stride(0), 0, M, N, K, None, N, None, 0, None, 0,
                                   tile_n, splits, L.ptr(sl), L.stream_ptr(a.device))
    L.check(rc, "b200surv_gemm_bf16_ex")
    assert (sl.sum(0) - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
