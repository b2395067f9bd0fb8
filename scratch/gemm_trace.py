"""Phase stamps of the CTA-pair GEMM (CTA 0): where do the microseconds go?"""
import torch
from multimodal_survival_prediction_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda", 0); L.require_device(0)
B, K, H, Kp = 4096, 5005, 512, 5008
x = torch.randn(B, Kp, device=dev).to(torch.bfloat16)[:, :K]
w = torch.randn(H, Kp, device=dev).to(torch.bfloat16)[:, :K]
dh = torch.randn(B, H, device=dev).to(torch.bfloat16)
st = L.stream_ptr(dev)
tr = torch.zeros(64, dtype=torch.int64, device=dev)
lib.b200surv_debug_gemm_trace(L.ptr(tr))
def show(tag):
    torch.cuda.synchronize()
    t = tr.cpu().tolist()
    rel = [(v - t[0]) / 1e3 for v in t]
    print(f"{tag}: first-mma-wait {rel[1]:.2f}  last-commit {rel[2]:.2f}  epi-begin {rel[3]:.2f}  epi-end {rel[4]:.2f}  exit {rel[5]:.2f} us")
    kb = [rel[8 + i] for i in range(24)]
    print("   k-block arrivals: " + " ".join(f"{v:.2f}" for v in kb))
    print("   deltas          : " + " ".join(f"{b - a:.2f}" for a, b in zip(kb[:-1], kb[1:])))
for warm in (0, 1):
    for splits in (1, 2):
        c = torch.empty(splits, B, H, device=dev)
        for rep in range(2):
            assert lib.b200surv_gemm_bf16_ex(L.ptr(x), x.stride(0), 0, L.ptr(w), w.stride(0), 0, B, H, K, L.ptr(c), H, None, 0, None, 0,
                                             512, splits, L.ptr(c) if splits > 1 else None, st) == 0
        show(f"fwd pair splits {splits} (second of two launches)")
    dw = torch.empty(1, H, K, device=dev)
    for rep in range(2):
        assert lib.b200surv_gemm_bf16_ex(L.ptr(dh), H, 1, L.ptr(x), x.stride(0), 1, H, K, B, L.ptr(dw), K, None, 0, None, 0, 512, 1, None, st) == 0
    show("wgrad pair")
