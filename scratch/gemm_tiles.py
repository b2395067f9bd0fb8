"""Timing of the two rna_encoder.0 GEMMs at B = 4096 for the tile widths / K splits of b200surv_gemm_bf16_ex."""
import torch
from multimodal_survival_prediction_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda", 0); L.require_device(0)
B, K, H = 4096, 5005, 512
Kp = 5008
x = torch.randn(B, Kp, device=dev).to(torch.bfloat16)[:, :K]
w = torch.randn(H, Kp, device=dev).to(torch.bfloat16)[:, :K]
dh = torch.randn(B, H, device=dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

st = L.stream_ptr(dev)
flops = 2.0 * B * K * H
for tile_n, splits in ((128, 1), (256, 2), (512, 1), (512, 2)):
    c = torch.empty(max(splits, 1), B, H, device=dev)
    def fwd():
        rc = lib.b200surv_gemm_bf16_ex(L.ptr(x), x.stride(0), 0, L.ptr(w), w.stride(0), 0, B, H, K, L.ptr(c), H, None, 0, None, 0,
                                       tile_n, splits, L.ptr(c) if splits > 1 else None, st)
        assert rc == 0, lib.b200surv_last_error()
    t = timed(fwd)
    print(f"fwd   x[4096x5005] W[512x5005]  tile_n {tile_n} splits {splits}: {t:6.1f} us  {flops / t / 1e6:7.1f} TFLOP/s", flush=True)
for tile_n, splits in ((128, 1), (256, 1), (512, 1), (512, 2)):
    dw = torch.empty(max(splits, 1), H, K, device=dev)
    def wg():
        rc = lib.b200surv_gemm_bf16_ex(L.ptr(dh), H, 1, L.ptr(x), x.stride(0), 1, H, K, B, L.ptr(dw), K, None, 0, None, 0,
                                       tile_n, splits, L.ptr(dw) if splits > 1 else None, st)
        assert rc == 0, lib.b200surv_last_error()
    t = timed(wg)
    print(f"wgrad dH[4096x512]^T x[4096x5005] tile_n {tile_n} splits {splits}: {t:6.1f} us  {flops / t / 1e6:7.1f} TFLOP/s", flush=True)
