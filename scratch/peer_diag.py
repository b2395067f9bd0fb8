"""torchrun diagnostic: forward-only timings of the fused single-GPU kernel vs the peer-exchange kernel vs NCCL."""
import ctypes
import os
import time

import torch
import torch.distributed as dist

from multimodal_survival_prediction_b200 import _lib as L
from multimodal_survival_prediction_b200 import dist as bd
from multimodal_survival_prediction_b200 import synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 24
lh, ev, t = synth.cohort(n, 7 + rank)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
grad = torch.empty(n, dtype=torch.float32, device=dev)
lib = L.load()


def timeit(name, fn, iters=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"rank {rank} {name}: gpu {a.elapsed_time(b) / iters * 1e3:.1f} us/iter, host enqueue {(h1 - h0) / iters * 1e6:.1f} us/iter",
          flush=True)
    dist.barrier()


single = bd.ShardedCoxBinned(n, dev, exchange="nccl")
st = L.stream_ptr(dev)


def f_single():
    rc = lib.b200surv_cox_fwd(L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, 2, 0, L.COX_BINNED, 4096, ctypes.c_float(0.0),
                              L.ptr(single.loss), L.ptr(single.state), single.sb, L.ptr(single.ws), single.wb, st)
    assert rc == 0


timeit("single fused fwd", f_single)
timeit("single fwd+bwd", lambda: (f_single(), single.backward(x, tt, e, grad)))
if world == 1:   # the peer kernel against its own buffer: phase trace of the single-GPU forward
    peer = bd.ShardedCoxBinned(n, dev, exchange="nccl")
    peer.peers = bd.PeerBuffers(lib.b200surv_cox_peer_buffer_bytes(4096))

    def f_peer1():
        peer.epoch += 1
        rc = lib.b200surv_cox_binned_fwd_peer(L.ptr(x), L.ptr(tt), L.ptr(e), n, 2, 0, 4096, ctypes.c_float(0.0),
                                              L.ptr(peer.loss), L.ptr(peer.state), peer.sb, L.ptr(peer.ws), peer.wb,
                                              peer.peers.array, 1, 0, peer.epoch, st)
        assert rc == 0
    peer.forward = lambda *a: f_peer1()
else:
    peer = bd.ShardedCoxBinned(n, dev, exchange="peer")
timeit("peer fwd", lambda: peer.forward(x, tt, e))
NAMES = ["start", "pass1", "sync1", "reduced", "sync2(peer)", "flag_sent", "pulled", "lookback1", "terms", "lookback2", "end", "rehearsed"]


def show_trace(tag, ws):
    if not os.environ.get("B200SURV_PEER_TRACE"):
        return
    off = lib.b200surv_cox_peer_trace_offset(n, 4096)
    torch.cuda.synchronize()
    tr = ws[off:off + 96].view(torch.int64).cpu().tolist()
    print(f"rank {rank} {tag} trace (us since start): " + ", ".join(f"{nm} {(v - tr[0]) / 1e3:.1f}" for nm, v in zip(NAMES, tr)),
          flush=True)


show_trace("peer", peer.ws)
f_single()
show_trace("single", single.ws)
timeit("peer fwd+bwd", lambda: (peer.forward(x, tt, e), peer.backward(x, tt, e, grad)))
timeit("nccl fwd", lambda: single.forward(x, tt, e))
timeit("nccl fwd+bwd", lambda: (single.forward(x, tt, e), single.backward(x, tt, e, grad)))
small = torch.zeros(12292, dtype=torch.int64, device=dev)
timeit("bare all_reduce 98KB", lambda: dist.all_reduce(small))
dist.destroy_process_group()
