"""C-index at 1M patients (and 100k): algo 1 (pair counting) against algo 2 (sorted column tiles), same six counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import cindex as gci, synth
dev = torch.device("cuda", 0)
for n in (100_000, 1 << 20, 1 << 22):
    lh, ev, t = synth.cohort(n, 1234)
    x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
    res = {}
    for algo in (1, 2):
        if algo == 1 and n > (1 << 20):
            continue
        for _ in range(2):
            c = gci.cindex_counts(x, e, tt, 1e-8, algo=algo)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            c = gci.cindex_counts(x, e, tt, 1e-8, algo=algo)
        b.record(); torch.cuda.synchronize()
        res[algo] = (a.elapsed_time(b) / reps, c.cpu().tolist())
        print(f"n={n} algo {algo}: {res[algo][0]:.3f} ms  counts {res[algo][1]}")
    if 1 in res:
        assert res[1][1] == res[2][1]
