"""Where the e2e step's 3.3 ms go: pinned H2D copies alone, the public-API step, pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import cox as gcox, synth
dev = torch.device("cuda", 0)
n = 1 << 24
lh, ev, t = synth.cohort(n, 1234)
pin = [x.pin_memory() for x in (lh, ev, t)]

def timeit(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {dt * 1e3:.3f} ms")
    return dt

def copies():
    xd = pin[0].to(dev, non_blocking=True); ed = pin[1].to(dev, non_blocking=True); td = pin[2].to(dev, non_blocking=True)
    return xd, ed, td
dt = timeit("3 pinned H2D copies (151 MB)", copies)
print(f"  -> {150994944 / dt / 1e9:.1f} GB/s")
big = torch.empty(150994944, dtype=torch.uint8).pin_memory()
dbig = torch.empty(150994944, dtype=torch.uint8, device=dev)
dt = timeit("one 151 MB pinned H2D copy into a preallocated buffer", lambda: dbig.copy_(big, non_blocking=True))
print(f"  -> {150994944 / dt / 1e9:.1f} GB/s")

def step():
    xd = pin[0].to(dev, non_blocking=True).requires_grad_(True)
    ed = pin[1].to(dev, non_blocking=True)
    td = pin[2].to(dev, non_blocking=True)
    loss = gcox.neg_partial_log_likelihood(xd, ed, td)
    loss.backward()
    return float(loss.item()), xd.grad
timeit("e2e step (public API, auto mode)", step)
xd, ed, td = copies()
def compute_only():
    x = xd.detach().requires_grad_(True)
    loss = gcox.neg_partial_log_likelihood(x, ed, td)
    loss.backward()
    return float(loss.item())
timeit("public API fwd+bwd+item on resident inputs", compute_only)
def step_binned():
    xd = pin[0].to(dev, non_blocking=True).requires_grad_(True)
    ed = pin[1].to(dev, non_blocking=True)
    td = pin[2].to(dev, non_blocking=True)
    loss = gcox.neg_partial_log_likelihood(xd, ed, td, mode="binned")
    loss.backward()
    return float(loss.item()), xd.grad
timeit("e2e step, mode=binned", step_binned)
