import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_survival_prediction_b200 import cindex as gci, synth
dev = torch.device("cuda", 0)
n = 1 << 20
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, c.tolist()


print("full            : %.2f ms" % timed(lambda: gci.cindex_counts(x, e, tt, 1e-8))[0])
for w in (2, 4, 8):
    print(f"rows [0,n/{w})     : %.2f ms   tile shard 0/{w}: %.2f ms   tile shard {w-1}/{w}: %.2f ms" % (
        timed(lambda: gci.cindex_counts(x, e, tt, 1e-8, 0, n // w, 1))[0],
        timed(lambda: gci.cindex_counts_shard(x, e, tt, 0, w, 1e-8))[0],
        timed(lambda: gci.cindex_counts_shard(x, e, tt, w - 1, w, 1e-8))[0]))
