"""SORTED Cox fwd+bwd timing (few-ties cohort, continuous times)."""
import sys, torch
from multimodal_survival_prediction_b200 import synth, cox as gcox, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
lh, ev, t = synth.cohort(n, 1234, few_ties=True)
x, e, tt = lh.cuda(), ev.cuda(), t.cuda()
one = torch.ones(1, device="cuda")
def step():
    loss, state = gcox.cox_fwd_raw(x, tt, e, None, 1, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_SORTED, 0)
    return loss, gcox.cox_bwd_raw(one, state, x, tt, e, None, 1, L.COX_SORTED, 0)
for _ in range(3): step()
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
reps = 10
tf = tb = 0.0
for _ in range(reps):
    a.record(); loss, state = gcox.cox_fwd_raw(x, tt, e, None, 1, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_SORTED, 0); b.record()
    g = gcox.cox_bwd_raw(one, state, x, tt, e, None, 1, L.COX_SORTED, 0); c.record(); torch.cuda.synchronize()
    tf += a.elapsed_time(b); tb += b.elapsed_time(c)
print(f"n={n}: SORTED fwd {tf / reps:.3f} ms, bwd {tb / reps:.3f} ms, loss {loss.item():.6f}")
