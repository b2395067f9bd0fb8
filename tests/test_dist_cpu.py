"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in multimodal_survival_prediction_b200/dist.py:
shard bounds, which collective runs with which reduce op, and that sharded counts add up to the
single-process result.  The per-shard compute is injected (the CPU oracle) -- the CUDA kernels themselves are
covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_survival_prediction_b200 import dist as bd
from multimodal_survival_prediction_b200 import synth


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [bd.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_q):
    from oracle import cindex as oci
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lh, ev, t = synth.cohort(n, 3, risk_tie_frac=0.1)

    def count_fn(est, event, time, tol, a, b, algo):     # stands in for the CUDA kernel on this rank's rows
        return torch.from_numpy(oci.counts_brute(est.numpy(), event.numpy(), time.numpy(), tol, a, b))

    counts = bd.cindex_counts_sharded(lh, ev, t, 1e-8, 1, None, _count_fn=count_fn)
    # Cox: the exchange is a SUM all-reduce of integer per-bin aggregates + a MAX all-reduce
    a, b = bd.shard_bounds(n, rank, world)
    bins = torch.zeros(3 * 64 + 4, dtype=torch.int64)
    idx = t[a:b].long().clamp(max=63)
    bins.index_add_(0, idx, torch.ones(b - a, dtype=torch.int64))
    mx = torch.tensor([lh[a:b].max().item(), -1.0])
    dist.all_reduce(bins, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    out_q.put((rank, counts.tolist(), bins.tolist(), mx.tolist()))
    dist.destroy_process_group()


def test_sharded_cindex_and_cox_exchange_world2():
    from oracle import cindex as oci
    n, world = 3000, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lh, ev, t = synth.cohort(n, 3, risk_tie_frac=0.1)
    full = oci.counts_brute(lh.numpy(), ev.numpy(), t.numpy(), 1e-8).tolist()
    ref_bins = torch.zeros(3 * 64 + 4, dtype=torch.int64)
    ref_bins.index_add_(0, t.long().clamp(max=63), torch.ones(n, dtype=torch.int64))
    for rank, counts, bins, mx in res:
        assert counts == full                      # every rank ends with the global counters, bit-exact
        assert bins == ref_bins.tolist()           # integer aggregates: exact for any sharding
        assert abs(mx[0] - lh.max().item()) < 1e-12


def _a2a_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    from multimodal_survival_prediction_b200 import dist as bd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # host logic of the sample sort in front of the SORTED shard phases: pooled samples -> identical splitters on every
        # rank; rows routed by destination; the uneven all-to-all (gloo has none: the fallback path)
        n = 1000 + 37 * rank
        g = torch.Generator().manual_seed(50 + rank)
        t = torch.empty(n).exponential_(1.0 / 1000.0, generator=g)
        pooled = [torch.empty(256) for _ in range(world)]
        dist.all_gather(pooled, t[:256].contiguous())
        spl = bd.choose_splitters(torch.cat(pooled), world)
        dest = np.searchsorted(spl.numpy(), t.numpy(), side="right")
        perm = np.argsort(dest, kind="stable")
        in_splits = np.bincount(dest, minlength=world).tolist()
        all_splits = [None] * world
        dist.all_gather_object(all_splits, in_splits)
        out_splits = [all_splits[src][rank] for src in range(world)]
        recv = torch.empty(sum(out_splits))
        bd._all_to_all(recv, t[perm].contiguous(), out_splits, in_splits)
        lo = -np.inf if rank == 0 else float(spl[rank - 1])
        hi = np.inf if rank == world - 1 else float(spl[rank])
        ok = bool(((recv.numpy() >= lo) & (recv.numpy() < hi)).all())
        sums = [None] * world
        dist.all_gather_object(sums, (float(recv.double().sum()), float(t.double().sum()), recv.numel(), n))
        q.put((rank, ok, spl.tolist(), sums))
    finally:
        dist.destroy_process_group()


def test_sample_sort_host_logic_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert got[0][2] == got[1][2] and len(got[0][2]) == 1            # one splitter, the same on both ranks
    assert got[0][1] and got[1][1]                                    # every received time lies in the rank's range
    sums = got[0][3]
    assert abs(sum(x[0] for x in sums) - sum(x[1] for x in sums)) < 1e-6 * sum(x[1] for x in sums)   # nothing lost
    assert sum(x[2] for x in sums) == sum(x[3] for x in sums)
