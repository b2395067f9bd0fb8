"""GPU parity tests of the Cox loss: CUDA path (through the C ABI) vs the float64 oracle.

Tolerances (north_star: 1e-5 relative, fp32): loss |d| <= 1e-5 * max(1, |ref|);
gradient max|d| <= 1e-5 * max|ref| (+1e-9 absolute)."""
import numpy as np
import pytest
import torch

import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import _lib as L
from multimodal_survival_prediction_b200 import cox as gcox
from multimodal_survival_prediction_b200 import synth
from oracle import cox as ocox

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-5


def run_gpu(eta, ev, t, ties="efron", reduction="mean", mode="auto", efron_mean_over="event_times", nbins=None,
            gscale=1.0):
    x = torch.as_tensor(eta, dtype=torch.float32).cuda().requires_grad_(True)
    loss = pkg.neg_partial_log_likelihood(x, torch.as_tensor(ev).bool().cuda(), torch.as_tensor(t, dtype=torch.float32).cuda(),
                                          ties, reduction, efron_mean_over=efron_mean_over, mode=mode, nbins=nbins)
    (loss * gscale).backward()
    return float(loss), x.grad.cpu().numpy().astype(np.float64)


def grad_atol(ev, reduction="mean"):
    """Absolute slack for gradients that are exactly 0 in exact arithmetic (e.g. the last row of a
    tie-free cohort): 1e-6 of the largest possible |d loss / d log_hz| = 1/normaliser."""
    return 1e-6 if reduction == "sum" else 1e-6 / max(1.0, float(np.sqrt(np.asarray(ev).sum())))


def check(eta, ev, t, ties="efron", reduction="mean", mode="auto", efron_mean_over="event_times", nbins=None,
          loss_rtol=LOSS_RTOL, grad_rtol=GRAD_RTOL):
    eta32 = np.asarray(eta, np.float32)
    ref_l, ref_g = ocox.cox_nll(eta32.astype(np.float64), ev, np.asarray(t, np.float32), ties, reduction,
                                efron_mean_over=efron_mean_over)
    l, g = run_gpu(eta32, ev, t, ties, reduction, mode, efron_mean_over, nbins)
    assert abs(l - ref_l) <= loss_rtol * max(1.0, abs(ref_l)), (l, ref_l, mode, ties, reduction)
    scale = max(np.abs(ref_g).max(), 1e-30)
    err = np.abs(g - ref_g).max()
    assert err <= grad_rtol * scale + grad_atol(ev, reduction), (err, scale, mode, ties, reduction)
    return l, g


def tied_cohort(n, seed, tmax=50, p_event=0.4):
    rng = np.random.default_rng(seed)
    t = rng.integers(1, tmax + 1, n).astype(np.float32)
    ev = rng.random(n) < p_event
    eta = rng.normal(size=n).astype(np.float32)
    return eta, ev, t


@pytest.mark.parametrize("mode", ["small", "binned", "sorted"])
def test_golden_reference_fallback_vectors(golden, mode):
    """Tie-free golden vectors produced by the reference's own fallback loss."""
    g = golden("cox_fallback.npz")
    for c in g["cases"]:
        eta, ev, t = g[f"{c}/log_hz"], g[f"{c}/event"], g[f"{c}/time"]
        l, gr = run_gpu(eta, ev, t, mode=mode, efron_mean_over="events")
        ref_l = float(g[f"{c}/rnaseq_only/f64/loss"]); ref_g = g[f"{c}/rnaseq_only/f64/grad"]
        assert abs(l - ref_l) <= LOSS_RTOL * max(1.0, abs(ref_l)), (c, mode)
        assert np.abs(gr - ref_g).max() <= GRAD_RTOL * max(np.abs(ref_g).max(), 1e-30) + grad_atol(ev), (c, mode)


@pytest.mark.parametrize("ties", ["efron", "breslow"])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
@pytest.mark.parametrize("mode", ["small", "binned", "sorted"])
def test_tied_times_all_modes(ties, reduction, mode):
    for n, seed, tmax in ((1, 0, 3), (2, 1, 1), (3, 2, 2), (8, 3, 4), (100, 4, 10), (777, 5, 30), (2048, 6, 200)):
        eta, ev, t = tied_cohort(n, seed, tmax)
        if not ev.any():
            ev[0] = True
        check(eta, ev, t, ties, reduction, mode)


@pytest.mark.parametrize("mode", ["binned", "sorted"])
@pytest.mark.parametrize("ties", ["efron", "breslow"])
def test_large_cohorts(mode, ties):
    for n, seed in ((2049, 10), (5000, 11), (100_003, 12), (1 << 20, 13)):
        lh, ev, t = synth.cohort(n, seed)
        check(lh.numpy(), ev.numpy(), t.numpy(), ties, "mean", mode)


def test_efron_mean_over_events_option():
    eta, ev, t = tied_cohort(5000, 20, 40)
    check(eta, ev, t, "efron", "mean", "binned", efron_mean_over="events")
    check(eta, ev, t, "efron", "mean", "small" if len(t) <= 2048 else "sorted", efron_mean_over="events")


def test_known_answers():
    # KA2: all rows tied, all events
    rng = np.random.default_rng(3)
    eta = rng.normal(size=700).astype(np.float32); ev = np.ones(700, bool); t = np.full(700, 4.0, np.float32)
    for mode in ("small", "binned", "sorted"):
        for ties in ("efron", "breslow"):
            check(eta, ev, t, ties, "sum", mode)
    # single event at the very end / very beginning
    ev2 = np.zeros(700, bool); ev2[13] = True
    t2 = np.arange(700, dtype=np.float32)
    for mode in ("small", "binned", "sorted"):
        check(eta, ev2, t2, "efron", "mean", mode)


def test_no_events_gives_zero_loss_and_zero_grad():
    eta, ev, t = tied_cohort(3000, 7)
    ev[:] = False
    for mode in ("small", "binned", "sorted"):
        n = 1000 if mode == "small" else 3000
        l, g = run_gpu(eta[:n], ev[:n], t[:n], mode=mode)
        assert l == 0.0 and np.all(g == 0.0)


def test_auto_mode_policy_and_headers():
    # small
    eta, ev, t = tied_cohort(100, 1)
    x = torch.tensor(eta).cuda()
    loss, state = gcox.cox_fwd_raw(x, torch.tensor(t).cuda(), torch.tensor(ev).cuda(), None, 1, L.TIES["efron"], 0,
                                   L.COX_SMALL, 0)
    h = gcox.read_headers(state, 1)[0]
    assert h.mode == L.COX_SMALL and h.flags == 0 and h.n_events == int(ev.sum())
    assert h.n_event_times == len(np.unique(t[ev]))
    # integer days beyond 4096 -> auto escalates to 8192 bins
    lh, ev, t = synth.cohort(50_000, 3)
    t = t * 2.0
    check(lh.numpy(), ev.numpy(), t.numpy())
    # non-integer times -> NOT_BINNABLE flag in binned mode (NaN loss), auto falls through to sorted
    lh, ev, t = synth.cohort(50_000, 4, few_ties=True)
    l, _ = run_gpu(lh.numpy(), ev.numpy(), t.numpy(), mode="binned")
    assert np.isnan(l)
    check(lh.numpy(), ev.numpy(), t.numpy())
    # huge log-hazards -> EXP_RANGE rescale inside auto; result equals the shifted problem
    lh, ev, t = synth.cohort(20_000, 5)
    check((lh + 100.0).numpy(), ev.numpy(), t.numpy())
    check((lh - 70.0).numpy(), ev.numpy(), t.numpy())


def test_bad_times_raise():
    lh, ev, t = synth.cohort(5000, 6)
    t[17] = -1.0
    with pytest.raises(ValueError):
        run_gpu(lh.numpy(), ev.numpy(), t.numpy())


def test_unaligned_views_and_dtypes():
    lh, ev, t = synth.cohort(10_007, 8)
    x = lh.cuda()[3:].requires_grad_(True)          # 12-byte offset: scalar path
    e = ev.cuda()[3:]; tt = t.cuda()[3:]
    loss = pkg.neg_partial_log_likelihood(x, e, tt, mode="binned")
    loss.backward()
    ref_l, ref_g = ocox.cox_nll(lh[3:].numpy().astype(np.float64), ev[3:].numpy(), t[3:].numpy())
    assert abs(float(loss) - ref_l) <= LOSS_RTOL * abs(ref_l)
    assert np.abs(x.grad.cpu().numpy() - ref_g).max() <= GRAD_RTOL * np.abs(ref_g).max()
    # float64 log_hz of shape (n,1) on the CPU: computed on the GPU, gradient returned on the CPU in float64
    x64 = lh[:500].double().reshape(-1, 1).requires_grad_(True)
    loss = pkg.neg_partial_log_likelihood(x64, ev[:500], t[:500])
    assert loss.device.type == "cpu" and loss.dim() == 0
    (2.5 * loss).backward()
    ref_l, ref_g = ocox.cox_nll(lh[:500].numpy().astype(np.float64), ev[:500].numpy(), t[:500].numpy())
    assert x64.grad.shape == (500, 1) and x64.grad.dtype == torch.float64
    assert np.abs(x64.grad.numpy()[:, 0] - 2.5 * ref_g).max() <= GRAD_RTOL * 2.5 * np.abs(ref_g).max()


@pytest.mark.parametrize("n", [4099, 65_537, 300_002])
def test_sorted_mode_unaligned_views_equal_aligned_run(n):
    """SORTED mode: the key kernel takes four rows per thread when the vectors are 16-byte aligned and a scalar loop for the
    rest (misaligned views, the n mod 4 tail).  Views at 4 / 12-byte (1 / 3-byte for the event vector) offsets must give the
    aligned run's loss (to one fp32 ulp: the fp64 sums meet through atomics) and the oracle's values to the usual tolerance; few-ties and heavy-ties cohorts."""
    for few in (True, False):
        lh, ev, t = synth.cohort(n + 3, 40 + n % 7, few_ties=few)
        outs = []
        for off in (0, 1, 3):
            base = [v.cuda() for v in (lh, ev, t)]
            # a fresh buffer whose element `off` holds row 0 of the cohort: same rows, different alignment
            bx = torch.empty(n + 8, device="cuda"); bt = torch.empty(n + 8, device="cuda")
            be = torch.empty(n + 8, dtype=torch.bool, device="cuda")
            bx[off:off + n] = base[0][:n]; bt[off:off + n] = base[2][:n]; be[off:off + n] = base[1][:n]
            x = bx[off:off + n].requires_grad_(True)
            loss = pkg.neg_partial_log_likelihood(x, be[off:off + n], bt[off:off + n], mode="sorted")
            loss.backward()
            outs.append((loss.detach().cpu(), x.grad.cpu()))
        for l, g in outs[1:]:
            assert abs(float(l) - float(outs[0][0])) <= 2e-7 * abs(float(outs[0][0]))    # fp64 sums by atomics: one fp32 ulp
            assert torch.allclose(g, outs[0][1], rtol=0, atol=1e-7 * float(outs[0][1].abs().max()))
        ref_l, ref_g = ocox.cox_nll(lh[:n].numpy().astype(np.float64), ev[:n].numpy(), t[:n].numpy())
        assert abs(float(outs[0][0]) - ref_l) <= LOSS_RTOL * abs(ref_l)
        assert np.abs(outs[0][1].numpy() - ref_g).max() <= GRAD_RTOL * np.abs(ref_g).max()


def test_sorted_beyond_one_scan_round_per_cluster_cta_equals_binned():
    """More than 8 x 1024 tiles (n > 16.7M rows): every CTA of the tile-scan clusters walks several rounds and carries between
    them.  No CPU oracle at this size; the check is the BINNED path, which shares no kernel with SORTED (integer days, Efron
    and Breslow): loss and gradient agree to the usual tolerance.  Also packed cohorts at a size where the packed tile count
    passes 8192."""
    n = 20_000_003
    lh, ev, t = synth.cohort(n, 77)
    e, tt = ev.cuda(), t.cuda()
    for ties in ("efron", "breslow"):
        out = {}
        for mode in ("binned", "sorted"):
            x = lh.cuda().requires_grad_(True)
            loss = pkg.neg_partial_log_likelihood(x, e, tt, ties_method=ties, mode=mode)
            loss.backward()
            out[mode] = (float(loss), x.grad)
        assert abs(out["sorted"][0] - out["binned"][0]) <= LOSS_RTOL * abs(out["binned"][0]), (ties, out["sorted"][0], out["binned"][0])
        gmax = float(out["binned"][1].abs().max())
        assert float((out["sorted"][1] - out["binned"][1]).abs().max()) <= GRAD_RTOL * gmax, ties
    del out
    # packed cohorts: 3 x 5.6M rows -> 8,205 + 3 tiles
    lens = [5_600_001, 5_600_002, 5_600_003]
    off = np.concatenate([[0], np.cumsum(lens)])
    m = int(off[-1])
    losses = {}
    for mode in ("binned", "sorted"):
        x = lh[:m].cuda().requires_grad_(True)
        ls = pkg.neg_partial_log_likelihood_segmented(x, ev[:m].cuda(), t[:m].cuda(), torch.tensor(off), mode=mode)
        ls.sum().backward()
        losses[mode] = (ls.detach().cpu().numpy(), x.grad)
    np.testing.assert_allclose(losses["sorted"][0], losses["binned"][0], rtol=LOSS_RTOL)
    gmax = float(losses["binned"][1].abs().max())
    assert float((losses["sorted"][1] - losses["binned"][1]).abs().max()) <= GRAD_RTOL * gmax


@pytest.mark.parametrize("mode", ["small", "binned"])
def test_segmented_cohorts(mode):
    rng = np.random.default_rng(9)
    if mode == "small":
        lens = [1, 2, 7, 2048, 100, 333, 5]
    else:
        lens = [5001, 2, 12_345, 4096, 40_003]
    off = np.concatenate([[0], np.cumsum(lens)])
    n = int(off[-1])
    lh, ev, t = synth.cohort(n, 21)
    t = torch.clamp(torch.floor(t / 40.0), 1, 4000)   # heavier ties
    x = lh.cuda().requires_grad_(True)
    w = torch.linspace(0.5, 2.0, len(lens)).cuda()
    losses = pkg.neg_partial_log_likelihood_segmented(x, ev.cuda(), t.cuda(), torch.tensor(off), mode=mode)
    (losses * w).sum().backward()
    ref_l, ref_g = ocox.cox_nll_segmented(lh.numpy().astype(np.float64), ev.numpy(), t.numpy(), off)
    np.testing.assert_allclose(losses.detach().cpu().numpy(), ref_l, rtol=LOSS_RTOL, atol=1e-6)
    g = x.grad.cpu().numpy()
    for s in range(len(lens)):
        a, b = off[s], off[s + 1]
        rg = ref_g[a:b] * float(w[s])
        assert np.abs(g[a:b] - rg).max() <= GRAD_RTOL * max(np.abs(rg).max(), 1e-30) + 2e-6, s


def test_full_size_16m_properties_and_oracle():
    """BASELINE.json configs[2]: 16,777,216 patients, ~30 % events, heavy ties, Efron."""
    n = 1 << 24
    lh, ev, t = synth.cohort(n, 1234)
    x = lh.cuda().requires_grad_(True); e = ev.cuda(); tt = t.cuda()
    loss = pkg.neg_partial_log_likelihood(x, e, tt, "efron", "sum", mode="binned")
    loss.backward()
    g = x.grad.double()
    # property: sum of the gradient of the summed loss is zero; loss invariant to a shift of log_hz
    assert abs(float(g.sum())) <= 1e-6 * float(g.abs().sum())
    loss_shift = pkg.neg_partial_log_likelihood((x.detach() + 1.75), e, tt, "efron", "sum", mode="binned")
    assert abs(float(loss_shift) - float(loss)) <= 2e-6 * abs(float(loss))
    # two independent GPU algorithms agree
    x2 = lh.cuda().requires_grad_(True)
    loss2 = pkg.neg_partial_log_likelihood(x2, e, tt, "efron", "sum", mode="sorted")
    loss2.backward()
    assert abs(float(loss2) - float(loss)) <= 2e-6 * abs(float(loss))
    assert float((x2.grad - x.grad).abs().max()) <= GRAD_RTOL * float(x.grad.abs().max())
    # and the float64 oracle at full size
    ref_l, ref_g = ocox.cox_nll(lh.numpy().astype(np.float64), ev.numpy(), t.numpy(), "efron", "sum")
    assert abs(float(loss) - ref_l) <= LOSS_RTOL * abs(ref_l)
    assert np.abs(x.grad.cpu().numpy() - ref_g).max() <= GRAD_RTOL * np.abs(ref_g).max()


def test_segmented_auto_mode_reads_every_segments_header():
    """BINNED interleaves the per-segment headers with the (P,F) tables; auto mode must read the header of
    EVERY segment (cohorts > 2048 rows, the CV-sweep case), including one that needs the EXP_RANGE retry and, in a
    second cohort set, one with a negative time."""
    lens = [5001, 3000, 12_345, 4096]
    off = np.concatenate([[0], np.cumsum(lens)])
    lh, ev, t = synth.cohort(int(off[-1]), 31)
    t = torch.clamp(torch.floor(t / 40.0), 1, 4000)
    lh = lh.clone()
    lh[off[2]:off[3]] += 60.0                      # only the third cohort is out of the fixed-point range at shift 0
    x = lh.cuda().requires_grad_(True)
    losses = pkg.neg_partial_log_likelihood_segmented(x, ev.cuda(), t.cuda(), torch.tensor(off))      # mode="auto"
    losses.sum().backward()
    ref_l, ref_g = ocox.cox_nll_segmented(lh.numpy().astype(np.float64), ev.numpy(), t.numpy(), off)
    np.testing.assert_allclose(losses.detach().cpu().numpy(), ref_l, rtol=LOSS_RTOL, atol=1e-6)
    g = x.grad.cpu().numpy()
    for s in range(len(lens)):
        a, b = off[s], off[s + 1]
        assert np.abs(g[a:b] - ref_g[a:b]).max() <= GRAD_RTOL * np.abs(ref_g[a:b]).max() + 2e-6, s
    # the headers themselves, at their real (strided) positions
    so = torch.tensor(off).cuda()
    _, state = gcox.cox_fwd_raw(lh.cuda(), t.cuda(), ev.cuda(), so, len(lens), L.TIES["efron"], 0, L.COX_BINNED, 4096)
    hdrs = gcox.read_headers(state, len(lens), L.COX_BINNED)
    assert [h.mode for h in hdrs] == [L.COX_BINNED] * 4 and [h.nbins for h in hdrs] == [4096] * 4
    assert [bool(h.flags & L.COXF_EXP_RANGE) for h in hdrs] == [False, False, True, False]
    for s, h in enumerate(hdrs):
        a, b = off[s], off[s + 1]
        assert h.n_events == int(ev[a:b].sum()) and h.max_log_hz == float(lh[a:b].max()) and h.min_log_hz == float(lh[a:b].min())
    t_bad = t.clone(); t_bad[off[3] + 5] = -2.0    # a bad time in the LAST cohort only
    with pytest.raises(ValueError):
        pkg.neg_partial_log_likelihood_segmented(lh.cuda(), ev.cuda(), t_bad.cuda(), torch.tensor(off))


def test_wide_hazard_spread_leaves_the_fixed_point_path():
    """BINNED quantises exp(log_hz - shift) to 2^-28; rows ~20 nats below the shift round to zero.  A late risk
    set made only of such rows used to give log(0).  Now the header raises LOW_PRECISION (loss NaN in explicit binned
    mode) and auto re-shifts or falls through to the fp64 SORTED path."""
    n = 20_000
    lh, ev, t = synth.cohort(n, 41)
    lh = lh.clone()
    late = t >= torch.quantile(t, 0.97)
    lh[late] -= 32.0                               # the largest times only hold very low hazards
    ev = ev.clone(); ev[late] = True
    l, _ = run_gpu(lh.numpy(), ev.numpy(), t.numpy(), mode="binned")
    assert np.isnan(l)
    x = lh.cuda()
    _, state = gcox.cox_fwd_raw(x, t.cuda(), ev.cuda(), None, 1, L.TIES["efron"], 0, L.COX_BINNED, 4096)
    h = gcox.read_headers(state, 1, L.COX_BINNED)[0]
    assert h.flags & L.COXF_LOW_PRECISION and h.min_log_hz == float(lh.min())
    check(lh.numpy(), ev.numpy(), t.numpy())       # auto: spread of ~37 nats does not fit any shift -> SORTED
    # a spread that fits after re-shifting stays BINNED (12 nats below; n = 20k allows weights up to 2^14) ...
    lh2, ev2, t2 = synth.cohort(n, 42)
    late2 = t2 >= torch.quantile(t2, 0.97)
    lh2 = lh2.clone(); lh2[late2] -= 12.0
    check(lh2.numpy(), ev2.numpy(), t2.numpy())
    # ... and two more nats do not: SORTED again, whose tie-group sums must not be differences of running totals
    lh2[late2] -= 2.0
    check(lh2.numpy(), ev2.numpy(), t2.numpy())
    for spread in (25.0, 60.0):                    # explicit SORTED mode on heavy ties + very wide spreads
        lh3, ev3, t3 = synth.cohort(30_000, 43)
        lh3 = lh3.clone(); lh3[t3 >= torch.quantile(t3, 0.9)] -= spread
        check(lh3.numpy(), ev3.numpy(), t3.numpy(), mode="sorted")


def test_small_cohort_bad_times_raise_like_large_ones():
    lh, ev, t = synth.cohort(40, 6)
    t[7] = float("nan")
    with pytest.raises(ValueError):
        run_gpu(lh.numpy(), ev.numpy(), t.numpy())
    t[7] = -3.0
    with pytest.raises(ValueError):
        run_gpu(lh.numpy(), ev.numpy(), t.numpy())
    # checks=False keeps the call asynchronous (CUDA-graph capture): the loss is NaN instead
    x = lh.cuda()
    assert torch.isnan(pkg.neg_partial_log_likelihood(x, ev.cuda(), t.cuda(), checks=False))


def test_sorted_mode_packed_cohorts_float_times_and_ties():
    """SORTED with cohorts packed back to back (the CV-sweep shape with continuous times, which used to raise): ragged cohort
    sizes around the 2048-row scan tiles and 4096-key sort tiles, float times, integer days with heavy ties, a cohort without
    events, cohorts whose hazards sit at very different levels; explicit mode and mode="auto"."""
    lens = [5001, 1, 4096, 2049, 12_345, 2, 30_000, 777]
    off = np.concatenate([[0], np.cumsum(lens)])
    n = int(off[-1])
    lh, ev, t = synth.cohort(n, 51, few_ties=True)
    lh, ev, t = lh.clone(), ev.clone(), t.clone()
    t[off[2]:off[3]] = torch.floor(t[off[2]:off[3]] / 50.0)          # one cohort with integer times and heavy ties
    t[off[6]:off[7]] = torch.round(t[off[6]:off[7]] * 4) / 4          # quarter-day ties
    ev[off[3]:off[4]] = False                                         # a cohort without events: loss 0, zero gradient
    lh[off[4]:off[5]] += 40.0                                         # levels far apart: every cohort has its own shift
    lh[off[6]:off[7]] -= 25.0
    w = torch.linspace(0.5, 2.0, len(lens)).cuda()
    for mode in ("sorted", "auto"):
        x = lh.cuda().requires_grad_(True)
        losses = pkg.neg_partial_log_likelihood_segmented(x, ev.cuda(), t.cuda(), torch.tensor(off), mode=mode)
        (losses * w).sum().backward()
        ref_l, ref_g = ocox.cox_nll_segmented(lh.numpy().astype(np.float64), ev.numpy(), t.numpy(), off)
        np.testing.assert_allclose(losses.detach().cpu().numpy(), ref_l, rtol=LOSS_RTOL, atol=1e-6, err_msg=mode)
        g = x.grad.cpu().numpy()
        for s in range(len(lens)):
            a, b = off[s], off[s + 1]
            rg = ref_g[a:b] * float(w[s])
            assert np.abs(g[a:b] - rg).max() <= GRAD_RTOL * max(np.abs(rg).max(), 1e-30) + 2e-6, (mode, s)
    # breslow and "sum" through the same path
    for ties, red in (("breslow", "mean"), ("efron", "sum")):
        losses = pkg.neg_partial_log_likelihood_segmented(lh.cuda(), ev.cuda(), t.cuda(), torch.tensor(off), ties, red, mode="sorted")
        ref_l, _ = ocox.cox_nll_segmented(lh.numpy().astype(np.float64), ev.numpy(), t.numpy(), off, ties_method=ties, reduction=red)
        np.testing.assert_allclose(losses.cpu().numpy(), ref_l, rtol=LOSS_RTOL, atol=1e-6)
    # many small cohorts (two sort passes on the cohort id: more than 256 cohorts)
    lens2 = [37] * 300
    off2 = np.concatenate([[0], np.cumsum(lens2)])
    lh2, ev2, t2 = synth.cohort(int(off2[-1]), 52, few_ties=True)
    losses = pkg.neg_partial_log_likelihood_segmented(lh2.cuda(), ev2.cuda(), t2.cuda(), torch.tensor(off2), mode="sorted")
    ref_l, _ = ocox.cox_nll_segmented(lh2.numpy().astype(np.float64), ev2.numpy(), t2.numpy(), off2)
    np.testing.assert_allclose(losses.cpu().numpy(), ref_l, rtol=LOSS_RTOL, atol=1e-6)
