"""The hand-written device primitives behind the SORTED Cox path and the C-index preprocessing (csrc/sortscan.cuh):
stable LSD radix sort and single-pass scans with decoupled look-back, against torch on sizes around the tile
boundaries (2048-element scan tiles, 4096-key sort tiles) and at millions of elements."""
import pytest
import torch

from multimodal_survival_prediction_b200 import _lib as L

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 2047, 2048, 2049, 4095, 4096, 4097, 8192 + 5, 100_003, 3_000_017]


def _temp(lib, n, dev):
    return torch.empty(lib.b200surv_debug_sortscan_temp_bytes(n), dtype=torch.uint8, device=dev)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("few_distinct", [False, True, "max"])
def test_radix_sort_pairs_is_a_stable_sort(n, few_distinct):
    dev = torch.device("cuda", 0)
    L.require_device(0)
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(n)
    hi = 7 if few_distinct is True else (1 << 32)       # many ties exercise stability
    keys = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64)
    if few_distinct == "max":
        # keys 0xffffffff and 0xfffffffe only: a partial tile pads itself with 0xffffffff keys, which must stay behind the real
        # ones and out of the published digit counts of every pass
        keys = (1 << 32) - 1 - (keys & 1)
    if few_distinct is False and n > 4:
        keys[:3] = torch.tensor([0, (1 << 32) - 1, 1 << 31])
    vals = torch.arange(n, dtype=torch.int64)
    kd = (keys & 0xFFFFFFFF).to(torch.int64).to(dev)
    # device buffers as int32 bit patterns
    kbuf = torch.empty(n, dtype=torch.int32, device=dev)
    kbuf.copy_(torch.where(kd >= (1 << 31), kd - (1 << 32), kd).to(torch.int32))
    vbuf = vals.to(torch.int32).to(dev)
    kt, vt = torch.empty_like(kbuf), torch.empty_like(vbuf)
    tmp = _temp(lib, n, dev)
    L.check(lib.b200surv_debug_sort_pairs(L.ptr(kbuf), L.ptr(vbuf), L.ptr(kt), L.ptr(vt), n, L.ptr(tmp), L.stream_ptr(dev)),
            "sort")
    torch.cuda.synchronize()
    got_k = kbuf.to(torch.int64).cpu() & 0xFFFFFFFF
    got_v = vbuf.to(torch.int64).cpu()
    ref_k, ref_v = torch.sort(keys, stable=True)
    assert torch.equal(got_k, ref_k)
    assert torch.equal(got_v, ref_v)           # stable: equal keys keep their input order


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("iop,reverse", [(0, 0), (0, 1), (1, 1), (2, 0)])
def test_lookback_scan_matches_torch(n, iop, reverse):
    dev = torch.device("cuda", 0)
    L.require_device(0)
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(7 * n + iop)
    a = torch.rand(n, generator=g, dtype=torch.float64)
    i = torch.randint(-1000, 1000, (n,), generator=g, dtype=torch.int64)
    ad, idv = a.to(dev), i.to(dev)
    oa, ob, oi = torch.empty_like(ad), torch.empty_like(ad), torch.empty_like(idv)
    tmp = _temp(lib, n, dev)
    L.check(lib.b200surv_debug_scan(L.ptr(ad), L.ptr(idv), n, iop, reverse, L.ptr(oa), L.ptr(ob), L.ptr(oi), L.ptr(tmp),
                                    L.stream_ptr(dev)), "scan")
    torch.cuda.synchronize()
    af, iff = (a.flip(0), i.flip(0)) if reverse else (a, i)
    ra = torch.cumsum(af, 0)
    ri = torch.cumsum(iff, 0) if iop == 0 else (torch.cummin(iff, 0).values if iop == 1 else torch.cummax(iff, 0).values)
    if reverse:
        ra, ri = ra.flip(0), ri.flip(0)
    assert torch.equal(oi.cpu(), ri)
    torch.testing.assert_close(oa.cpu(), ra, rtol=1e-12, atol=1e-9)
    torch.testing.assert_close(ob.cpu(), 2 * ra, rtol=1e-12, atol=1e-9)
