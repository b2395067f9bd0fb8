"""SORTED Cox loss over TIME-RANGE shards (SURVEY.md 8e path "Cox (B)"; BASELINE.json north_star: all-gather of per-shard
boundary aggregates with exact carry-in): several shards driven by ONE process on one GPU -- the arithmetic that crosses
shards is exactly what `dist.ShardedCoxSorted` exchanges over NCCL (three 128-byte records per shard), here concatenated
by hand -- against the float64 oracle on the whole cohort and against the single-GPU SORTED path.

What the reference does on one device: argsort(time) + logcumsumexp (scripts/training/partial_modality_training.py:303-309).
Tolerances: loss 1e-5 * max(1, |ref|), gradient 1e-5 * max|ref| (north_star), vs the single-GPU kernels 2e-6."""
import numpy as np
import pytest
import torch

import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import _lib as L
from multimodal_survival_prediction_b200 import dist as bd
from multimodal_survival_prediction_b200 import synth
from oracle import cox as ocox

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-5


def run_shards(eta, ev, t, cuts, ties="efron", reduction=L.REDUCE_MEAN_TERMS, shuffle_seed=0):
    """Rows sorted by time are cut at `cuts` into shards (each shard's rows then shuffled: a shard is an unordered set of
    rows of one time range); returns loss, gradient in the ORIGINAL row order, and the header of shard 0."""
    dev = torch.device("cuda", 0)
    order = np.argsort(np.asarray(t), kind="stable")
    bounds = [0] + list(cuts) + [len(order)]
    rng = np.random.default_rng(shuffle_seed)
    rows = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        r = order[a:b].copy()
        rng.shuffle(r)
        rows.append(r)
    world = len(rows)
    ops, data = [], []
    for r, idx in enumerate(rows):
        ops.append(bd.ShardedCoxSorted(len(idx), dev, ties="efron" if ties == "efron" else "breslow", reduction=reduction,
                                       rank=r, world=world))
        data.append((torch.as_tensor(np.asarray(eta, np.float32)[idx]).to(dev), torch.as_tensor(np.asarray(t, np.float32)[idx]).to(dev),
                     torch.as_tensor(np.asarray(ev, bool)[idx]).to(dev)))
    rec0 = torch.cat([op.phase_keys(*d).clone() for op, d in zip(ops, data)])
    for op in ops:
        op.phase_sort()
    rec1 = torch.cat([op.phase_reduce(d[0], rec0).clone() for op, d in zip(ops, data)])
    rec2 = torch.cat([op.phase_terms(rec1).clone() for op in ops])
    losses = [float(op.phase_finish(rec2)) for op in ops]
    grad = np.zeros(len(order), np.float64)
    for op, idx in zip(ops, rows):
        g = torch.empty(len(idx), dtype=torch.float32, device=dev)
        op.backward(g)
        grad[idx] = g.cpu().numpy()
    assert all(l == losses[0] or (np.isnan(l) and np.isnan(losses[0])) for l in losses), losses   # one value on every shard
    return losses[0], grad, ops[0].header(), ops


def check(eta, ev, t, cuts, ties="efron", reduction="mean"):
    red = {"mean": L.REDUCE_MEAN_TERMS, "sum": L.REDUCE_SUM}[reduction]
    eta32 = np.asarray(eta, np.float32)
    ref_l, ref_g = ocox.cox_nll(eta32.astype(np.float64), ev, np.asarray(t, np.float32), ties, reduction)
    l, g, hdr, _ = run_shards(eta32, ev, t, cuts, ties, red)
    assert hdr.flags == 0
    assert abs(l - ref_l) <= LOSS_RTOL * max(1.0, abs(ref_l)), (l, ref_l, cuts)
    err, scale = np.abs(g - ref_g).max(), max(np.abs(ref_g).max(), 1e-30)
    assert err <= GRAD_RTOL * scale + 1e-9, (err, scale, cuts)
    return l, g, hdr


def tied(n, seed, tmax, p_event=0.4):
    rng = np.random.default_rng(seed)
    return rng.normal(size=n).astype(np.float32), rng.random(n) < p_event, rng.integers(1, tmax + 1, n).astype(np.float32)


@pytest.mark.parametrize("ties", ["efron", "breslow"])
def test_tie_groups_across_shard_edges(ties):
    """Heavy ties: every cut falls inside a tie group; groups span two, three and all shards; groups whose events sit on
    one side of an edge only (the head of the group is a censored row of the previous shard)."""
    eta, ev, t = tied(9000, 3, tmax=6)
    check(eta, ev, t, [1234, 4000, 4001, 7777], ties)          # ~1500 rows per tie group: cuts inside groups, a 1-row shard
    check(eta, ev, t, [2048, 4096, 6144], ties)                # cuts on tile boundaries
    eta, ev, t = tied(5000, 4, tmax=1)                         # ONE tie group over every shard
    check(eta, ev, t, [100, 2148, 2149, 4500], ties)
    # events on one side of an edge only: time 2's censored rows end shard 0 (it holds no event of time 2), its events sit
    # on shard 1 -- the distinct event time must still be counted once and its Efron terms start at l = 0
    t = np.array([1, 1, 2, 2, 2, 2, 3, 3], np.float32)
    ev = np.array([1, 0, 0, 0, 1, 1, 1, 0], bool)
    eta = np.linspace(-1, 1, 8).astype(np.float32)
    for cut in ([4], [3], [5], [2, 4, 6], [1, 2, 3, 4, 5, 6, 7]):
        _, _, hdr = check(eta, ev, t, cut, ties)
        assert hdr.n_events == 4 and hdr.n_event_times == 3


def test_shards_match_single_gpu_sorted_and_oracle_continuous_times():
    n = 300_000
    lh, ev, t = synth.cohort(n, 77, few_ties=True)
    cuts = [n // 8 * k + (7 * k) % 13 for k in range(1, 8)]    # eight ragged shards
    l, g, hdr = check(lh.numpy(), ev.numpy(), t.numpy(), cuts)
    x = lh.cuda().requires_grad_(True)
    l1 = pkg.neg_partial_log_likelihood(x, ev.cuda(), t.cuda(), "efron", mode="sorted")
    l1.backward()
    assert abs(l - float(l1)) <= 2e-6 * abs(float(l1))
    assert np.abs(g - x.grad.cpu().numpy()).max() <= 2e-6 * float(x.grad.abs().max())
    assert hdr.n_events == int(ev.sum())


def test_shards_heavy_ties_integer_days_sum_reduction():
    n = 200_000
    lh, ev, t = synth.cohort(n, 78)                            # integer days, ~50 rows per day
    check(lh.numpy(), ev.numpy(), t.numpy(), [50_000, 100_001, 150_002], "efron", "sum")
    check(lh.numpy(), ev.numpy(), t.numpy(), [199_999], "breslow", "mean")


def test_wide_hazard_spread_common_shift():
    """The exponent shift is the GLOBAL max log_hz (first record): a shard of low hazards must not use its own."""
    n = 40_000
    lh, ev, t = synth.cohort(n, 79, few_ties=True)
    lh = lh.clone()
    lh[t > t.median()] -= 25.0
    check(lh.numpy(), ev.numpy(), t.numpy(), [n // 2, n // 2 + 5000])


def test_shards_out_of_time_order_are_flagged():
    dev = torch.device("cuda", 0)
    lh, ev, t = synth.cohort(4000, 80, few_ties=True)
    ops = [bd.ShardedCoxSorted(2000, dev, rank=r, world=2) for r in range(2)]
    data = [(lh[a:a + 2000].to(dev), t[a:a + 2000].to(dev), ev[a:a + 2000].to(dev)) for a in (0, 2000)]   # row blocks, not time ranges
    rec0 = torch.cat([op.phase_keys(*d).clone() for op, d in zip(ops, data)])
    for op in ops:
        op.phase_sort()
    rec1 = torch.cat([op.phase_reduce(d[0], rec0).clone() for op, d in zip(ops, data)])
    rec2 = torch.cat([op.phase_terms(rec1).clone() for op in ops])
    for op in ops:
        assert np.isnan(float(op.phase_finish(rec2)))
        assert op.header().flags & L.COXF_NOT_PARTITIONED
        with pytest.raises(L.B200SurvError):
            op.check()


def test_one_shard_equals_the_plain_sorted_path():
    lh, ev, t = synth.cohort(50_000, 81, few_ties=True)
    dev = torch.device("cuda", 0)
    op = bd.ShardedCoxSorted(50_000, dev, rank=0, world=1)
    loss = op.forward(lh.to(dev), t.to(dev), ev.to(dev))
    g = torch.empty(50_000, dtype=torch.float32, device=dev)
    op.backward(g)
    x = lh.cuda().requires_grad_(True)
    l1 = pkg.neg_partial_log_likelihood(x, ev.cuda(), t.cuda(), "efron", mode="sorted")
    l1.backward()
    assert float(loss) == float(l1) and torch.equal(g, x.grad)   # same kernels, same tiles: bit-identical


def test_route_rows_groups_by_destination_stably():
    """b200surv_route_rows: destination = number of splitters <= time, rows grouped by destination in their original order,
    per-destination counts, packed send buffers; gather / scatter through the permutation are inverse to each other."""
    dev = torch.device("cuda", 0)
    lib = L.load()
    for n, n_dest in ((1, 1), (5, 3), (10_000, 8), (300_001, 64)):
        lh, ev, t = synth.cohort(n, 90 + n_dest)
        spl = np.sort(np.random.default_rng(n).choice(np.unique(t.numpy()), size=min(n_dest - 1, len(np.unique(t.numpy()))), replace=False)).astype(np.float32)
        spl = np.concatenate([spl, np.full(n_dest - 1 - len(spl), np.inf, np.float32)])
        x, tt, e = lh.to(dev), t.to(dev), ev.to(dev)
        ws = torch.empty(lib.b200surv_route_workspace_bytes(n), dtype=torch.uint8, device=dev)
        o_lh, o_t = torch.empty(n, device=dev), torch.empty(n, device=dev)
        o_e = torch.empty(n, dtype=torch.bool, device=dev)
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        counts = torch.empty(n_dest, dtype=torch.int64, device=dev)
        sp = torch.as_tensor(spl).to(dev)
        L.check(lib.b200surv_route_rows(L.ptr(x), L.ptr(tt), L.ptr(e), n, L.ptr(sp) if n_dest > 1 else None, n_dest, L.ptr(o_lh),
                                        L.ptr(o_t), L.ptr(o_e), L.ptr(perm), L.ptr(counts), L.ptr(ws), ws.numel(), L.stream_ptr(dev)),
                "b200surv_route_rows")
        dest = np.searchsorted(spl, t.numpy(), side="right")
        want = np.argsort(dest, kind="stable")
        assert np.array_equal(perm.cpu().numpy(), want)
        assert np.array_equal(counts.cpu().numpy(), np.bincount(dest, minlength=n_dest))
        assert torch.equal(o_lh.cpu(), lh[want]) and torch.equal(o_t.cpu(), t[want]) and torch.equal(o_e.cpu(), ev[want])
        g = torch.empty(n, device=dev)
        L.check(lib.b200surv_route_gather(L.ptr(x), L.ptr(perm), n, L.ptr(g), L.stream_ptr(dev)), "gather")
        assert torch.equal(g, o_lh)
        back = torch.empty(n, device=dev)
        L.check(lib.b200surv_route_scatter(L.ptr(g), L.ptr(perm), n, L.ptr(back), L.stream_ptr(dev)), "scatter")
        assert torch.equal(back, x)


def test_row_block_operator_on_one_rank_equals_sorted():
    dev = torch.device("cuda", 0)
    n = 70_001
    lh, ev, t = synth.cohort(n, 95, few_ties=True)
    op = bd.RowBlockCoxSorted(n, dev).plan(t.to(dev), ev.to(dev))
    loss = op.forward(lh.to(dev))
    g = torch.empty(n, device=dev)
    op.backward(g)
    assert op.check() == 0
    x = lh.cuda().requires_grad_(True)
    l1 = pkg.neg_partial_log_likelihood(x, ev.cuda(), t.cuda(), "efron", mode="sorted")
    l1.backward()
    assert float(loss) == float(l1.detach()) and torch.equal(g, x.grad)
