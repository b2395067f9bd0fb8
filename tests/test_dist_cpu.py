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
