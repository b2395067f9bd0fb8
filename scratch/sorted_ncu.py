import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import synth
n = 1 << 24
lh, ev, t = synth.cohort(n, 3, few_ties=True)
x, e, tt = lh.cuda().requires_grad_(True), ev.cuda(), t.cuda()
for _ in range(2):
    loss = pkg.neg_partial_log_likelihood(x, e, tt, mode="sorted")
    loss.backward()
torch.cuda.synchronize()
