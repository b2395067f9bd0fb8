"""Row-block sharding of the hot path across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  The reference has no distributed code;
this is the scale-out of the two operators it calls:

* C-index: every rank holds the full (estimate, event, time) vectors (9 bytes/row), counts the pairs
  of its own row tiles (dealt out round-robin in sorted order) against all columns, then ONE int64 SUM all-reduce of the six counters
  (48 bytes) -- integer, so the result is bit-identical to the single-GPU result.
* Cox loss (SORTED mode, time-range shards): `ShardedCoxSorted` below.
* Cox loss (BINNED mode): every rank accumulates per-bin aggregates of its own rows
  (b200surv_cox_binned_partial; 32.32 fixed-point integers), ONE int64 SUM all-reduce of
  3*nbins+4 words (exact, so the loss is bit-identical for every sharding) plus a 2-float MAX
  all-reduce, then every rank finalises identically (b200surv_cox_binned_finalize) and computes the
  gradient of its own rows (b200surv_cox_bwd).  The gradient stays sharded like the input.

The host logic (shard bounds, which collectives with which reduce op) is backend-agnostic and is
covered on CPU with gloo at world_size 2 (tests/test_dist_cpu.py) by injecting the per-shard compute.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib as L


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous row block [a, b) of rank; blocks differ by at most one row."""
    base, rem = divmod(n, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ------------------------------------------------------------------ C-index
def cindex_counts_sharded(estimate, event, time, tied_tol=1e-8, algo=1, group=None, _count_fn=None):
    """All ranks pass the same full vectors; returns the global int64[6] counters on every rank.

    On the GPU rank r counts the row TILES r, r + world, ... of the sorted event rows (cindex_counts_shard): its
    kernel keeps the tile structure of the single-GPU run.  algo 1 (the default here): the pair-by-pair tile kernel the
    8-GPU scaling target is quoted on; algo 2: ranks in sorted tiles (25x faster on one GPU at 1M patients, so its shards
    are short against the replicated preprocessing).  A caller-supplied ``_count_fn(est, ev, time, tol, a, b,
    algo)`` (the CPU tests of this host logic) gets the contiguous row block ``shard_bounds(n, rank, world)``."""
    rank, world = _world()
    if _count_fn is None and algo in (1, 2):
        from .cindex import cindex_counts_shard
        counts = cindex_counts_shard(estimate, event, time, rank, world, tied_tol, algo=algo)
    else:
        if _count_fn is None:
            from .cindex import cindex_counts as _count_fn
        a, b = shard_bounds(estimate.numel(), rank, world)
        counts = _count_fn(estimate, event, time, tied_tol, a, b, algo)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


# ------------------------------------------------------------------ peer buffers (CUDA IPC over NVLink)
class PeerBuffers:
    """One zero-filled device buffer per rank, mapped into every rank of the box (b200surv_peer_alloc/open).
    The 64-byte IPC handles travel through the process group's object all-gather (host side, once)."""

    def __init__(self, nbytes: int, group=None):
        self.lib = L.load()
        self.rank, self.world = _world()
        self.error = None            # every rank runs every collective below, whatever fails locally
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        rc = self.lib.b200surv_peer_alloc(nbytes, ctypes.byref(own), handle)
        if rc != L.OK:
            self.error = "b200surv_peer_alloc: " + (self.lib.b200surv_last_error() or b"").decode()
        self.own = own.value if rc == L.OK else None
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if rc == L.OK else b"", group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
                continue
            p = ctypes.c_void_p()
            if len(h) != 64:
                self.error = self.error or f"rank {r} exported no handle"
            elif self.lib.b200surv_peer_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)) != L.OK:
                self.error = self.error or "b200surv_peer_open: " + (self.lib.b200surv_last_error() or b"").decode()
            self.ptrs.append(p.value)
        self.array = (ctypes.c_void_p * self.world)(*[x or 0 for x in self.ptrs])
        oks = [None] * self.world     # doubles as the barrier: every buffer is mapped and zero-filled before first use
        dist.all_gather_object(oks, self.error is None, group=group)
        if not all(oks) and self.error is None:
            self.error = "a peer rank could not map the buffers"

    def close(self):
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                self.lib.b200surv_peer_close(ctypes.c_void_p(p))
        if self.own:
            self.lib.b200surv_peer_free(ctypes.c_void_p(self.own))
        self.ptrs, self.own = [], None


# ------------------------------------------------------------------ Cox (BINNED)
class ShardedCoxBinned:
    """Pre-allocated buffers for repeated sharded fwd+bwd over this rank's rows.

    exchange: how the per-bin sums meet across ranks
      "peer" -- fused into the forward kernel over peer memory (b200surv_cox_binned_fwd_peer): ONE launch, no
                collective call; needs all ranks on one box (CUDA IPC);
      "nccl" -- partial -> one int64 SUM all-reduce (torch.distributed) -> finalize;
      "auto" -- "peer" when the peer buffers can be set up on every rank, else "nccl".
    Both give the bit-identical loss (integer sums)."""

    def __init__(self, n_local: int, device, nbins: int = 4096, ties: str = "efron", reduction: int = L.REDUCE_MEAN_TERMS,
                 sync_max: bool = False, exchange: str = "auto", group=None):
        self.lib = L.load()
        self.sync_max = sync_max
        self.peers, self.epoch, self.exchange = None, 0, "nccl"
        _, world = _world()
        if world > 1 and exchange in ("auto", "peer") and not sync_max:
            self.peers = PeerBuffers(self.lib.b200surv_cox_peer_buffer_bytes(nbins), group)
            if self.peers.error is None:       # the same verdict on every rank (PeerBuffers gathers it)
                self.exchange = "peer"
            else:
                err = self.peers.error
                self.peers.close()
                self.peers = None
                if exchange == "peer":
                    raise L.B200SurvError("peer exchange unavailable: " + err)
        L.require_device(device.index)
        self.n, self.nb, self.dev = n_local, nbins, device
        self.ties, self.red = L.TIES[ties], reduction
        self.cnt = self.lib.b200surv_cox_bins_sum_count(nbins)
        self.bins_sum = torch.empty(self.cnt, dtype=torch.int64, device=device)
        self.bins_max = torch.empty(2, dtype=torch.float32, device=device)
        self.sb = self.lib.b200surv_cox_state_bytes(n_local, 1, L.COX_BINNED, nbins)
        self.wb = self.lib.b200surv_cox_workspace_bytes(n_local, 1, L.COX_BINNED, nbins)
        self.state = torch.empty(self.sb, dtype=torch.uint8, device=device)
        self.ws = torch.empty(self.wb, dtype=torch.uint8, device=device)
        self.loss = torch.empty(1, dtype=torch.float32, device=device)
        self.ones = torch.ones(1, dtype=torch.float32, device=device)

    def forward(self, log_hz, time, event, shift: float = 0.0, group=None):
        st = L.stream_ptr(self.dev)
        _, world = _world()
        if world == 1:  # nothing to exchange: the fused single-GPU forward (one launch fewer)
            rc = self.lib.b200surv_cox_fwd(L.ptr(log_hz), L.ptr(time), L.ptr(event), None, self.n, 1, self.ties,
                                           self.red, L.COX_BINNED, self.nb, ctypes.c_float(shift), L.ptr(self.loss),
                                           L.ptr(self.state), self.sb, L.ptr(self.ws), self.wb, st)
            L.check(rc, "b200surv_cox_fwd")
            return self.loss
        if self.exchange == "peer":   # the exchange happens inside the kernel, over NVLink peer memory
            self.epoch += 1
            rc = self.lib.b200surv_cox_binned_fwd_peer(L.ptr(log_hz), L.ptr(time), L.ptr(event), self.n, self.ties,
                                                       self.red, self.nb, ctypes.c_float(shift), L.ptr(self.loss),
                                                       L.ptr(self.state), self.sb, L.ptr(self.ws), self.wb,
                                                       self.peers.array, self.peers.world, self.peers.rank,
                                                       self.epoch & 0xFFFFFFFF or 1, st)
            L.check(rc, "b200surv_cox_binned_fwd_peer")
            return self.loss
        rc = self.lib.b200surv_cox_binned_partial(L.ptr(log_hz), L.ptr(time), L.ptr(event), None, self.n, 1, self.nb,
                                                  ctypes.c_float(shift), L.ptr(self.bins_sum), L.ptr(self.bins_max),
                                                  L.ptr(self.ws), self.wb, st)
        L.check(rc, "b200surv_cox_binned_partial")
        _, world = _world()
        if world > 1:
            # one collective per step: the per-bin aggregates (exact int64 SUM).  bins_max stays rank-local: it only
            # feeds the EXP_RANGE check of the explicit `shift`, which every rank then evaluates on its own rows
            # (pass sync_max=True to all-reduce it as well, e.g. to pick a common shift for the next call).
            dist.all_reduce(self.bins_sum, op=dist.ReduceOp.SUM, group=group)
            if self.sync_max:
                dist.all_reduce(self.bins_max, op=dist.ReduceOp.MAX, group=group)
        rc = self.lib.b200surv_cox_binned_finalize(L.ptr(self.bins_sum), L.ptr(self.bins_max), self.n, 1, self.ties,
                                                   self.red, self.nb, ctypes.c_float(shift), L.ptr(self.loss),
                                                   L.ptr(self.state), self.sb, L.ptr(self.ws), self.wb, st)
        L.check(rc, "b200surv_cox_binned_finalize")
        return self.loss

    def backward(self, log_hz, time, event, out_grad, grad_out=None):
        g = self.ones if grad_out is None else grad_out
        rc = self.lib.b200surv_cox_bwd(L.ptr(g), L.ptr(self.state), self.sb, L.ptr(log_hz), L.ptr(time), L.ptr(event),
                                       None, self.n, 1, L.COX_BINNED, self.nb, L.ptr(out_grad), L.stream_ptr(self.dev))
        L.check(rc, "b200surv_cox_bwd")
        return out_grad


# ------------------------------------------------------------------ Cox (SORTED, time-range shards)
class ShardedCoxSorted:
    """SORTED Cox loss over TIME-RANGE shards (SURVEY.md 8e path "Cox (B)", the formulation BASELINE.json's north_star
    gives for several GPUs): rank r holds ``n_local`` rows in any order whose times do not exceed the times of rank
    r + 1 (equal times may sit on both sides of an edge).  Every rank sorts and scans its own rows; between the phases
    the ranks all-gather ONE 128-byte record each (``b200surv_cox_sorted_shard_*``, include/b200surv.h): max log_hz and
    edge times, then the shard's tile sequence folded into one element per scan chain, then the sums behind the loss.
    Every rank folds the records of the other ranks into its exact carry-in on the device; no host synchronisation.
    The first all-gather overlaps the radix sort.  The gradient stays sharded like the input.

    ``rank`` / ``world`` default to the process group's; passing them explicitly (with ``gather=`` a callable that maps
    this shard's record to the ``world`` records) lets one process drive several shards, which is how the single-GPU
    parity test runs the multi-shard arithmetic."""

    REC = 128

    def __init__(self, n_local: int, device, ties: str = "efron", reduction: int = L.REDUCE_MEAN_TERMS, rank=None,
                 world=None):
        self.lib = L.load()
        L.require_device(device.index)
        r, w = _world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        if self.world > 64:
            raise L.B200SurvError("ShardedCoxSorted: at most 64 shards")
        if n_local < 1:
            raise ValueError("every shard needs at least one row")
        assert self.lib.b200surv_cox_shard_record_bytes() == self.REC
        self.n, self.capacity, self.dev = n_local, n_local, device     # n may shrink per call (set_rows), never above capacity
        self.ties, self.red = L.TIES[ties], reduction
        self.sb = self.lib.b200surv_cox_state_bytes(n_local, 1, L.COX_SORTED, 0)
        self.wb = self.lib.b200surv_cox_workspace_bytes(n_local, 1, L.COX_SORTED, 0)
        self.state = torch.empty(self.sb, dtype=torch.uint8, device=device)
        self.ws = torch.empty(self.wb, dtype=torch.uint8, device=device)
        self.rec = torch.zeros(3, self.REC, dtype=torch.uint8, device=device)
        self.all = torch.zeros(3, self.world * self.REC, dtype=torch.uint8, device=device)
        self.loss = torch.empty(1, dtype=torch.float32, device=device)
        self.ones = torch.ones(1, dtype=torch.float32, device=device)

    def set_rows(self, n: int):
        """Rows of this shard for the next forward / backward pair (1 <= n <= the capacity given at construction): the
        buffers are sized once, a re-partitioned cohort may hand a rank a different number of rows."""
        if not 1 <= n <= self.capacity:
            raise ValueError(f"shard rows {n} outside [1, {self.capacity}]")
        self.n = n

    # -- the four phases (each returns the record the caller must all-gather before the next one)
    def phase_keys(self, log_hz, time, event):
        L.check(self.lib.b200surv_cox_sorted_shard_keys(L.ptr(log_hz), L.ptr(time), L.ptr(event), self.n, L.ptr(self.rec[0]),
                                                        L.ptr(self.ws), self.wb, L.stream_ptr(self.dev)),
                "b200surv_cox_sorted_shard_keys")
        return self.rec[0]

    def phase_sort(self):
        L.check(self.lib.b200surv_cox_sorted_shard_sort(self.n, L.ptr(self.ws), self.wb, L.stream_ptr(self.dev)),
                "b200surv_cox_sorted_shard_sort")

    def phase_reduce(self, log_hz, all_rec0):
        L.check(self.lib.b200surv_cox_sorted_shard_reduce(L.ptr(log_hz), self.n, L.ptr(all_rec0), self.rank, self.world,
                                                          L.ptr(self.rec[1]), L.ptr(self.ws), self.wb, L.stream_ptr(self.dev)),
                "b200surv_cox_sorted_shard_reduce")
        return self.rec[1]

    def phase_terms(self, all_rec1):
        L.check(self.lib.b200surv_cox_sorted_shard_terms(self.n, self.ties, L.ptr(all_rec1), self.rank, self.world,
                                                         L.ptr(self.rec[2]), L.ptr(self.ws), self.wb, L.stream_ptr(self.dev)),
                "b200surv_cox_sorted_shard_terms")
        return self.rec[2]

    def phase_finish(self, all_rec2):
        L.check(self.lib.b200surv_cox_sorted_shard_finish(self.n, self.ties, self.red, L.ptr(all_rec2), self.rank, self.world,
                                                          L.ptr(self.loss), L.ptr(self.state), self.sb, L.ptr(self.ws), self.wb,
                                                          L.stream_ptr(self.dev)),
                "b200surv_cox_sorted_shard_finish")
        return self.loss

    def forward(self, log_hz, time, event, group=None):
        """Loss of the whole cohort (the same value on every rank); the state keeps this rank's gradient."""
        rec0 = self.phase_keys(log_hz, time, event)
        if self.world == 1:
            self.all[0].copy_(rec0)
            self.phase_sort()
        else:
            work = dist.all_gather_into_tensor(self.all[0], rec0, group=group, async_op=True)
            self.phase_sort()              # does not need the records: the collective runs beside it
            work.wait()
        rec1 = self.phase_reduce(log_hz, self.all[0])
        if self.world == 1:
            self.all[1].copy_(rec1)
        else:
            dist.all_gather_into_tensor(self.all[1], rec1, group=group)
        rec2 = self.phase_terms(self.all[1])
        if self.world == 1:
            self.all[2].copy_(rec2)
        else:
            dist.all_gather_into_tensor(self.all[2], rec2, group=group)
        return self.phase_finish(self.all[2])

    def backward(self, out_grad, grad_out=None):
        g = self.ones if grad_out is None else grad_out
        rc = self.lib.b200surv_cox_bwd(L.ptr(g), L.ptr(self.state), self.sb, None, None, None, None, self.n, 1, L.COX_SORTED,
                                       0, L.ptr(out_grad), L.stream_ptr(self.dev))
        L.check(rc, "b200surv_cox_bwd")
        return out_grad

    def header(self) -> "L.CoxHeader":
        """The cohort's header (synchronises): flags, loss, number of events / distinct event times ..."""
        raw = self.state[:L.COX_HEADER_BYTES].cpu().numpy().tobytes()
        return L.CoxHeader.from_buffer_copy(raw)

    def check(self):
        """Raise like the single-GPU operator does: ValueError for bad times, B200SurvError for shards out of time order."""
        flags = self.header().flags
        if flags & L.COXF_BAD_TIME:
            raise ValueError("time must be finite and non-negative")
        if flags & L.COXF_NOT_PARTITIONED:
            raise L.B200SurvError("ShardedCoxSorted: the shards are not time ranges (a rank holds a time above the next "
                                  "rank's smallest time); use time_range_partition first")
        return flags


# ------------------------------------------------------------------ row blocks -> time ranges (sample sort) + SORTED shards
def choose_splitters(samples: torch.Tensor, world: int) -> torch.Tensor:
    """world - 1 ascending cut points of the time axis from the pooled samples of all ranks (the same tensor on every rank
    gives the same cuts): the k/world quantiles of the sorted sample."""
    s, _ = torch.sort(samples.flatten())
    idx = (torch.arange(1, world, device=s.device) * s.numel()) // world
    return s[idx].contiguous()


def _all_to_all(out, inp, out_splits, in_splits, group=None):
    """all_to_all_single with uneven splits; on backends without it (gloo: the CPU tests of this host logic) the same exchange
    as one broadcast per (source, destination) pair of a padded all-gather."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, list(in_splits), group=group)
    cap = max(sum(x) for x in sizes)
    pad = torch.zeros(cap, dtype=inp.dtype)
    pad[:inp.numel()] = inp
    bufs = [torch.zeros(cap, dtype=inp.dtype) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    o = 0
    for src in range(world):
        a = sum(sizes[src][:rank])
        k = sizes[src][rank]
        out[o:o + k] = bufs[src][a:a + k]
        o += k


class RowBlockCoxSorted:
    """SORTED Cox loss for patients sharded by ROW BLOCK over the ranks (BASELINE.json north_star): a sample sort on
    survival time puts every rank in charge of one time range, then `ShardedCoxSorted` (boundary records all-gathered,
    exact carry-in).  What the reference does with one argsort on one device (partial_modality_training.py:303-309).

    ``plan(time, event)`` -- once per cohort; times and events do not change between training steps: pooled samples ->
    world - 1 splitters, rank-local routing (b200surv_route_rows: destination of every row, rows grouped by destination,
    the permutation back), one all-to-all each for time and event.  Reads the per-destination counts on the host (the only
    synchronisation).  ``forward(log_hz)`` -- per step: pack through the kept permutation (b200surv_route_gather), one
    all-to-all of 4 bytes per row, the shard phases; returns the loss of the WHOLE cohort (same value on every rank).
    ``backward(out_grad)``: the shard's gradient rows travel back (all-to-all) into the caller's row order
    (b200surv_route_scatter)."""

    def __init__(self, n_local: int, device, ties: str = "efron", reduction: int = L.REDUCE_MEAN_TERMS, slack: float = 1.25,
                 samples_per_rank: int = 4096, group=None):
        self.lib = L.load()
        L.require_device(device.index)
        self.rank, self.world = _world()
        self.n, self.dev, self.group = n_local, device, group
        self.ties, self.red, self.slack, self.samples = ties, reduction, slack, samples_per_rank
        self.rwb = self.lib.b200surv_route_workspace_bytes(n_local)
        self.rws = torch.empty(self.rwb, dtype=torch.uint8, device=device)
        self.perm = torch.empty(n_local, dtype=torch.int32, device=device)
        self.counts = torch.zeros(max(self.world, 1), dtype=torch.int64, device=device)
        self.send = torch.empty(n_local, dtype=torch.float32, device=device)
        self.shard = None
        self.n_recv = 0

    def plan(self, time, event):
        n, dev, world = self.n, self.dev, self.world
        if world > 1:
            stride = max(1, n // self.samples)
            mine = time[::stride][:self.samples].contiguous()
            if mine.numel() < self.samples:        # short shards: repeat (weights the pooled quantiles by shard, roughly)
                mine = mine.repeat((self.samples + mine.numel() - 1) // mine.numel())[:self.samples].contiguous()
            pooled = torch.empty(world * self.samples, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(pooled, mine, group=self.group)
            self.splitters = choose_splitters(pooled, world)
        else:
            self.splitters = torch.empty(0, dtype=torch.float32, device=dev)
        t_send = torch.empty(n, dtype=torch.float32, device=dev)
        e_send = torch.empty(n, dtype=torch.bool, device=dev)
        L.check(self.lib.b200surv_route_rows(None, L.ptr(time), L.ptr(event), n, L.ptr(self.splitters) if world > 1 else None, world,
                                             None, L.ptr(t_send), L.ptr(e_send), L.ptr(self.perm), L.ptr(self.counts), L.ptr(self.rws),
                                             self.rwb, L.stream_ptr(dev)), "b200surv_route_rows")
        self.in_splits = [int(x) for x in self.counts.cpu().tolist()]                # the one host synchronisation
        if world > 1:
            recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_to_all_single(recv_counts, self.counts, group=self.group)
            self.out_splits = [int(x) for x in recv_counts.cpu().tolist()]
        else:
            self.out_splits = list(self.in_splits)
        self.n_recv = sum(self.out_splits)
        if self.n_recv < 1:
            raise L.B200SurvError("RowBlockCoxSorted: a rank received no rows (fewer distinct times than ranks?)")
        if self.shard is None or self.n_recv > self.shard.capacity:
            self.shard = ShardedCoxSorted(max(int(self.n_recv * self.slack), self.n_recv), dev, ties=self.ties, reduction=self.red)
            cap = self.shard.capacity
            self.t_recv = torch.empty(cap, dtype=torch.float32, device=dev)
            self.e_recv = torch.empty(cap, dtype=torch.bool, device=dev)
            self.x_recv = torch.empty(cap, dtype=torch.float32, device=dev)
            self.g_recv = torch.empty(cap, dtype=torch.float32, device=dev)
        self.shard.set_rows(self.n_recv)
        if world > 1:
            _all_to_all(self.t_recv[:self.n_recv], t_send, self.out_splits, self.in_splits, self.group)
            _all_to_all(self.e_recv[:self.n_recv].view(torch.uint8), e_send.view(torch.uint8), self.out_splits, self.in_splits, self.group)
        else:
            self.t_recv[:n].copy_(t_send); self.e_recv[:n].copy_(e_send)
        return self

    def forward(self, log_hz):
        if self.shard is None:
            raise L.B200SurvError("RowBlockCoxSorted.forward before plan(time, event)")
        L.check(self.lib.b200surv_route_gather(L.ptr(log_hz), L.ptr(self.perm), self.n, L.ptr(self.send), L.stream_ptr(self.dev)),
                "b200surv_route_gather")
        x = self.x_recv[:self.n_recv]
        if self.world > 1:
            _all_to_all(x, self.send, self.out_splits, self.in_splits, self.group)
        else:
            x.copy_(self.send)
        return self.shard.forward(x, self.t_recv[:self.n_recv], self.e_recv[:self.n_recv], group=self.group)

    def backward(self, out_grad, grad_out=None):
        g = self.g_recv[:self.n_recv]
        self.shard.backward(g, grad_out)
        if self.world > 1:
            _all_to_all(self.send, g, self.in_splits, self.out_splits, self.group)
        else:
            self.send.copy_(g)
        L.check(self.lib.b200surv_route_scatter(L.ptr(self.send), L.ptr(self.perm), self.n, L.ptr(out_grad), L.stream_ptr(self.dev)),
                "b200surv_route_scatter")
        return out_grad

    def check(self):
        return self.shard.check()
