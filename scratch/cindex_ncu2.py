import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import cindex as gci, synth
dev = torch.device("cuda", 0)
n = 1 << 20
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
for _ in range(2):
    c = gci.cindex_counts(x, e, tt, 1e-8, algo=int(os.environ.get("ALGO", "2")))
torch.cuda.synchronize()
