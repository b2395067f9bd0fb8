"""A/B of libb200surv builds on the two large GEMM shapes of the head (one process, interleaved)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_survival_prediction_b200 import _lib as L
dev = torch.device("cuda", 0)
paths = sys.argv[1:]
libs = []
for p in paths:
    lib = ctypes.CDLL(os.path.abspath(p))
    fn = lib.b200surv_gemm_bf16
    fn.restype, fn.argtypes = L.SIGNATURES["b200surv_gemm_bf16"]
    libs.append((os.path.basename(p), fn))
B, K, H = 4096, 5008, 512
x = torch.randn(B, K, device=dev).bfloat16(); w = torch.randn(H, K, device=dev).bfloat16()
dy = torch.randn(B, H, device=dev).bfloat16()
y = torch.empty(B, H, device=dev); dw = torch.empty(H, 5005, device=dev)
st = L.stream_ptr(dev)
def fwd(fn): assert fn(L.ptr(x), K, 0, L.ptr(w), K, 0, B, H, 5005, L.ptr(y), H, None, 0, None, 0, st) == 0
def wgrad(fn): assert fn(L.ptr(dy), H, 1, L.ptr(x), K, 1, H, 5005, B, L.ptr(dw), 5005, None, 0, None, 0, st) == 0
def timed(f, fn, it=30):
    for _ in range(3): f(fn)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f(fn)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
for r in range(3):
    print("  ".join(f"{nm}: fwd {timed(fwd, fn):.1f} us wgrad {timed(wgrad, fn):.1f} us" for nm, fn in libs), flush=True)
