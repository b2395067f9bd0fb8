"""One GPU: the fused forward as compiled for one GPU (PEER=false) against the peer variant run against its own buffer
(world 1), each inside the fwd, bwd, fwd, bwd, ... loop with CUDA events around every kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import _lib as L, dist as bd, synth
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n = 1 << 24
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
grad = torch.empty(n, dtype=torch.float32, device=dev)
lib = L.load()
st = L.stream_ptr(dev)
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29517")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
single = bd.ShardedCoxBinned(n, dev, exchange="nccl")
peer = bd.ShardedCoxBinned(n, dev, exchange="nccl")
peer.peers = bd.PeerBuffers(lib.b200surv_cox_peer_buffer_bytes(4096))

def f_single():
    rc = lib.b200surv_cox_fwd(L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, 2, 0, L.COX_BINNED, 4096, ctypes.c_float(0.0),
                              L.ptr(single.loss), L.ptr(single.state), single.sb, L.ptr(single.ws), single.wb, st)
    assert rc == 0
def f_peer():
    peer.epoch += 1
    rc = lib.b200surv_cox_binned_fwd_peer(L.ptr(x), L.ptr(tt), L.ptr(e), n, 2, 0, 4096, ctypes.c_float(0.0),
                                          L.ptr(peer.loss), L.ptr(peer.state), peer.sb, L.ptr(peer.ws), peer.wb,
                                          peer.peers.array, 1, 0, peer.epoch, st)
    assert rc == 0
def run(name, fwd, obj, reps=40):
    for _ in range(5):
        fwd(); obj.backward(x, tt, e, grad)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(reps):
        fwd(); obj.backward(x, tt, e, grad)
    a1.record()
    torch.cuda.synchronize()
    for a, b, c in evs:
        a.record(); fwd(); b.record(); obj.backward(x, tt, e, grad); c.record()
    torch.cuda.synchronize()
    f = sum(a.elapsed_time(b) for a, b, c in evs) / reps; bw = sum(b.elapsed_time(c) for a, b, c in evs) / reps
    print(f"{name}: loop {a0.elapsed_time(a1) / reps * 1e3:.1f} us/step; with events: fwd {f * 1e3:.1f} us, bwd {bw * 1e3:.1f} us", flush=True)
for rnd in range(3):
    run("single (PEER=false)", f_single, single)
    run("peer variant, world 1", f_peer, peer)
assert torch.equal(single.loss, peer.loss), (single.loss, peer.loss)
print("losses equal")
dist.destroy_process_group()
