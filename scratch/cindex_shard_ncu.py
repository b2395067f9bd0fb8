import os, sys, torch
sys.path.insert(0, os.getcwd())
from multimodal_survival_prediction_b200 import cindex as gci, synth
dev = torch.device("cuda", 0)
n = 1 << 20
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
for _ in range(3):
    gci.cindex_counts_shard(x, e, tt, 0, 8, 1e-8)
torch.cuda.synchronize()
