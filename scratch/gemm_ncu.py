"""One launch of each rna_encoder.0 GEMM variant, for `ncu --set full -k regex:gemm_bf16_tc`."""
import torch
from multimodal_survival_prediction_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda", 0); L.require_device(0)
B, K, H, Kp = 4096, 5005, 512, 5008
x = torch.randn(B, Kp, device=dev).to(torch.bfloat16)[:, :K]
w = torch.randn(H, Kp, device=dev).to(torch.bfloat16)[:, :K]
dh = torch.randn(B, H, device=dev).to(torch.bfloat16)
st = L.stream_ptr(dev)
for tile_n, splits in ((128, 1), (256, 2)):
    c = torch.empty(max(splits, 1), B, H, device=dev)
    assert lib.b200surv_gemm_bf16_ex(L.ptr(x), x.stride(0), 0, L.ptr(w), w.stride(0), 0, B, H, K, L.ptr(c), H, None, 0, None, 0,
                                     tile_n, splits, L.ptr(c) if splits > 1 else None, st) == 0
for tile_n, splits in ((128, 1), (256, 1)):
    dw = torch.empty(max(splits, 1), H, K, device=dev)
    assert lib.b200surv_gemm_bf16_ex(L.ptr(dh), H, 1, L.ptr(x), x.stride(0), 1, H, K, B, L.ptr(dw), K, None, 0, None, 0,
                                     tile_n, splits, L.ptr(dw) if splits > 1 else None, st) == 0
torch.cuda.synchronize()
print("ok")
