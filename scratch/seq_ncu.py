"""fwd,bwd,fwd,bwd ... for one library build, for `ncu --metrics gpu__time_duration.sum` (pure kernel durations, no launch gaps).
    PYTHONPATH=. python scratch/seq_ncu.py <lib.so> [pairs]"""
import ctypes, os, sys
import torch
from multimodal_survival_prediction_b200 import _lib as L
from multimodal_survival_prediction_b200 import synth
path, pairs = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 6
lib = ctypes.CDLL(os.path.abspath(path))
for name, (res, args) in L.SIGNATURES.items():
    if hasattr(lib, name):
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
n = 1 << 24
dev = torch.device("cuda", 0)
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
grad = torch.empty(n, dtype=torch.float32, device=dev)
sb = lib.b200surv_cox_state_bytes(n, 1, L.COX_BINNED, 4096); wb = lib.b200surv_cox_workspace_bytes(n, 1, L.COX_BINNED, 4096)
state = torch.zeros(sb, dtype=torch.uint8, device=dev); ws = torch.zeros(wb, dtype=torch.uint8, device=dev)
loss = torch.empty(1, dtype=torch.float32, device=dev); one = torch.ones(1, dtype=torch.float32, device=dev)
st = L.stream_ptr(dev)
for _ in range(pairs):
    assert lib.b200surv_cox_fwd(L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, 2, 0, L.COX_BINNED, 4096, ctypes.c_float(0.0), L.ptr(loss), L.ptr(state), sb, L.ptr(ws), wb, st) == 0
    assert lib.b200surv_cox_bwd(L.ptr(one), L.ptr(state), sb, L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, L.COX_BINNED, 4096, L.ptr(grad), st) == 0
torch.cuda.synchronize()
print(path, loss.item())
