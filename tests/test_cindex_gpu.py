"""GPU parity tests of the C-index pair counters: bit-exact int64 counts vs the C oracle."""
import numpy as np
import pytest
import torch

import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import synth
from oracle import cindex as oci

pytestmark = pytest.mark.gpu


def gpu_counts(est, ev, t, tol=1e-8, algo=1, row_begin=0, row_end=None):
    out = pkg.cindex_counts(torch.as_tensor(est, dtype=torch.float32).cuda(), torch.as_tensor(ev).bool().cuda(),
                            torch.as_tensor(t, dtype=torch.float32).cuda(), tol, row_begin, row_end, algo)
    return out.cpu().numpy()


def cohort(n, seed, tmax=7, risk_ties=False):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, tmax + 1, n).astype(np.float32)
    ev = rng.random(n) < 0.5
    est = rng.normal(size=n).astype(np.float32)
    if risk_ties:
        est = (np.round(est * 3) / 3).astype(np.float32)
    return est, ev, t


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_golden_reference_fallback_vectors(golden, algo):
    g = golden("cindex_fallback.npz")
    for c in g["cases"]:
        est, ev, t = g[f"{c}/est"], g[f"{c}/event"], g[f"{c}/time"]
        counts = gpu_counts(est, ev, t, 0.0, algo)
        assert (counts == oci.counts_brute(est, ev, t, 0.0)).all(), c
        assert np.float32(pkg.cindex_from_counts(counts, "fallback")) == g[f"{c}/value"], c


@pytest.mark.parametrize("algo", [0, 1, 2])
@pytest.mark.parametrize("tol", [0.0, 1e-8, 0.3])
def test_counts_bit_exact_small_and_ragged(algo, tol):
    for n, seed, tmax, rt in ((1, 0, 3, False), (2, 1, 1, False), (5, 2, 2, True), (100, 3, 7, True),
                              (1000, 4, 30, False), (2047, 5, 5, True), (2049, 6, 500, True), (5000, 7, 4000, False)):
        est, ev, t = cohort(n, seed, tmax, rt)
        assert (gpu_counts(est, ev, t, tol, algo) == oci.counts_brute(est, ev, t, tol)).all(), (n, seed)


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_known_answers_and_special_values(algo):
    est = np.zeros(64, np.float32); ev = np.ones(64, bool); t = np.arange(64, dtype=np.float32)
    c = gpu_counts(est, ev, t, 1e-8, algo)
    assert list(c) == [0, 0, 64 * 63 // 2, 0, 0, 0]                      # KA3
    c = gpu_counts(-t, ev, t, 1e-8, algo)
    assert list(c) == [64 * 63 // 2, 0, 0, 0, 0, 0]                      # KA4
    c = gpu_counts(est, ev, np.ones(64, np.float32), 1e-8, algo)
    assert c.sum() == 0                                                   # KA2
    # infinities and NaN estimates follow the fp32 predicate literally
    est, ev, t = cohort(300, 11, 9, True)
    est[::7] = np.inf; est[3::11] = -np.inf; est[5::13] = np.nan
    assert (gpu_counts(est, ev, t, 1e-8, algo) == oci.counts_brute(est, ev, t, 1e-8)).all()
    assert (gpu_counts(est, ev, t, 0.5, algo) == oci.counts_brute(est, ev, t, 0.5)).all()
    # times: zero, huge, infinite
    t2 = t.copy(); t2[::5] = 0.0; t2[1::9] = 3e38; t2[2::17] = np.inf
    assert (gpu_counts(est, ev, t2, 1e-8, algo) == oci.counts_brute(est, ev, t2, 1e-8)).all()


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_row_sharding_adds_up(algo):
    lh, ev, t = synth.cohort(30_000, 5, risk_tie_frac=0.1)
    full = oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())
    cuts = [0, 7000, 7001, 19_999, 30_000]
    acc = np.zeros(6, np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = gpu_counts(lh, ev, t, 1e-8, algo, a, b)
        assert (part == oci.counts_brute(lh.numpy(), ev.numpy(), t.numpy(), 1e-8, a, b)).all()
        acc += part
    assert (acc == full).all()


def test_moderate_and_full_size():
    for n, seed in ((100_000, 1), (1 << 20, 1234)):        # BASELINE.json configs[3]: 1M patients
        for frac in (0.0, 0.1):
            lh, ev, t = synth.cohort(n, seed, risk_tie_frac=frac)
            ref = oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())
            assert (gpu_counts(lh, ev, t, 1e-8, 1) == ref).all(), (n, frac)
            assert (gpu_counts(lh, ev, t, 1e-8, 2) == ref).all(), (n, frac, "sorted column tiles")
    lh, ev, t = synth.cohort(200_000, 2, risk_tie_frac=0.1)
    assert (gpu_counts(lh, ev, t, 1e-8, 0) == oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())).all()
    # few ties in time (float times)
    lh, ev, t = synth.cohort(100_000, 3, few_ties=True)
    assert (gpu_counts(lh, ev, t, 1e-8, 1) == oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())).all()
    assert (gpu_counts(lh, ev, t, 1e-8, 2) == oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())).all()


@pytest.mark.parametrize("tol", [0.0, 1e-8, 0.3])
def test_sorted_column_tiles_equal_pair_counting(tol):
    """algo 2 answers the strictly-later tiles with two binary searches per row in a tile sorted by estimate; the six
    integers must equal algo 1's pair-by-pair counts, also with infinite / NaN / signed-zero estimates, heavy risk ties
    and tie tolerances wider than the spacing of the estimates."""
    for n, seed, frac in ((60_000, 21, 0.0), (60_000, 22, 0.3), (131_072, 23, 0.05), (9_000, 24, 0.5)):
        lh, ev, t = synth.cohort(n, seed, risk_tie_frac=frac)
        est = lh.numpy().copy()
        est[::97] = np.inf; est[3::101] = -np.inf; est[5::103] = np.nan; est[7::107] = 0.0; est[11::109] = -0.0
        a = gpu_counts(est, ev, t, tol, 1)
        b = gpu_counts(est, ev, t, tol, 2)
        assert (a == b).all(), (n, seed, tol, a, b)
    lh, ev, t = synth.cohort(60_000, 25, few_ties=True)
    assert (gpu_counts(lh, ev, t, tol, 2) == gpu_counts(lh, ev, t, tol, 1)).all()


def test_concordance_index_object_api():
    lh, ev, t = synth.cohort(348, 0)
    ci = pkg.ConcordanceIndex()
    val = ci(lh, ev, t)                      # CPU tensors in, like the reference's validate()
    assert val.device.type == "cpu" and val.dtype == torch.float32 and val.dim() == 0
    ref = oci.counts_brute(lh.numpy(), ev.numpy(), t.numpy())
    assert ci.counts == list(ref)
    assert val.item() == np.float32(oci.cindex_from_counts(ref))
    val_gpu = pkg.ConcordanceIndex(convention="fallback")(lh.cuda(), ev.cuda(), t.cuda())
    assert val_gpu.is_cuda and val_gpu.item() == np.float32(oci.cindex_from_counts(ref, "fallback"))
    assert pkg.ConcordanceIndex()(lh[:0], ev[:0], t[:0]).item() == 0.5


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_packed_cohorts_match_per_cohort_counts(algo):
    """CV-sweep shape: ragged cohorts packed back to back (one empty), one launch sequence, counts per cohort."""
    from multimodal_survival_prediction_b200.cindex import cindex_counts_cohorts
    sizes = [1, 0, 37, 1000, 2049, 5]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    est, ev, t = cohort(int(offs[-1]), 11, tmax=9, risk_ties=True)
    out = cindex_counts_cohorts(torch.from_numpy(est).cuda(), torch.from_numpy(ev).cuda(), torch.from_numpy(t).cuda(),
                                offs.tolist(), 1e-8, algo).cpu().numpy()
    for c, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
        ref = oci.counts_brute(est[a:b], ev[a:b], t[a:b], 1e-8) if b > a else np.zeros(6, dtype=np.int64)
        assert (out[c] == ref).all(), (c, a, b)


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_tile_shards_partition_the_pairs(n_shards, algo):
    """b200surv_cindex_counts_shard / _shard_algo: the shards' counters add up to the single-GPU counters bit for bit (what
    the multi-GPU int64 all-reduce relies on), on a cohort with time ties, risk ties and several row tiles."""
    from multimodal_survival_prediction_b200 import synth
    from multimodal_survival_prediction_b200.cindex import cindex_counts, cindex_counts_shard
    lh, ev, t = synth.cohort(30_011 if algo == 1 else 90_011, 21, risk_tie_frac=0.1)
    x, e, tt = lh.cuda(), ev.cuda(), t.cuda()
    full = cindex_counts(x, e, tt, 1e-8, algo=1).cpu()
    acc = torch.zeros(6, dtype=torch.int64)
    for s in range(n_shards):
        part = cindex_counts_shard(x, e, tt, s, n_shards, 1e-8, algo=algo).cpu()
        assert (part >= 0).all()
        acc += part
    assert torch.equal(acc, full), (acc.tolist(), full.tolist())


@pytest.mark.parametrize("n", [1, 2, 37, 2048, 5000])
def test_tile_shards_on_small_cohorts_including_empty_shards(n):
    """Fewer row tiles than shards: the surplus shards count nothing, the sum is still exact."""
    from multimodal_survival_prediction_b200 import synth
    from multimodal_survival_prediction_b200.cindex import cindex_counts, cindex_counts_shard
    lh, ev, t = synth.cohort(n, 33, risk_tie_frac=0.2)
    t = torch.clamp(torch.floor(t / 200.0), 1, 40)          # many same-time pairs
    x, e, tt = lh.cuda(), ev.cuda(), t.cuda()
    full = cindex_counts(x, e, tt, 1e-8).cpu()
    ref = oci.counts_fast(lh.numpy(), ev.numpy(), t.numpy())
    assert full.tolist() == list(ref)
    acc = torch.zeros(6, dtype=torch.int64)
    for s in range(8):
        acc += cindex_counts_shard(x, e, tt, s, 8, 1e-8).cpu()
    assert torch.equal(acc, full)


def test_cohorts_in_flight_match_one_by_one():
    """b200surv_cindex_counts_cohorts with room for eight cohorts in flight (internal streams) equals per-cohort calls."""
    from multimodal_survival_prediction_b200 import synth
    from multimodal_survival_prediction_b200.cindex import cindex_counts, cindex_counts_cohorts
    lens = [3001, 1, 777, 20_000, 2, 4096, 9999, 12, 5000, 64, 8191]
    off = np.concatenate([[0], np.cumsum(lens)])
    lh, ev, t = synth.cohort(int(off[-1]), 17, risk_tie_frac=0.1)
    x, e, tt = lh.cuda(), ev.cuda(), t.cuda()
    out = cindex_counts_cohorts(x, e, tt, off.tolist()).cpu()
    for c in range(len(lens)):
        a, b = int(off[c]), int(off[c + 1])
        one = cindex_counts(x[a:b].contiguous(), e[a:b].contiguous(), tt[a:b].contiguous(), 1e-8).cpu()
        assert torch.equal(out[c], one), c


def test_negative_and_nan_times_raise():
    """algo 1's sort key would alias a negative time with its absolute value; torchsurv rejects such input."""
    lh, ev, t = synth.cohort(500, 3)
    for bad in (-1.0, float("nan")):
        t2 = t.clone(); t2[11] = bad
        with pytest.raises(ValueError):
            pkg.ConcordanceIndex()(lh, ev, t2)
        with pytest.raises(ValueError):
            pkg.ConcordanceIndex()(lh.cuda(), ev.cuda(), t2.cuda())
    assert 0.0 <= pkg.ConcordanceIndex()(lh, ev, t).item() <= 1.0
