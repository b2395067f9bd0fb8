"""BASELINE configs[0] step (CT encoder + ungated head + Cox loss + backward as one CUDA graph, clip + AdamW) at batch 4:
time per step; run with B200SURV_PDL=0 / 1 for the A/B of programmatic dependent launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import cox as gcox, head as ghead, synth
from multimodal_survival_prediction_b200.optim import ClipAdam
dev = torch.device("cuda", 0)
c1 = ghead.MultiModalSurvivalNet().to(dev).train()
opt = ClipAdam(c1.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=1.0, adamw=True)
_, rna, clin, _ = synth.modality_batch(4, seed=1234)
host = [torch.rand(4, 1, 64, 64, 32).pin_memory(), rna.pin_memory(), clin.pin_memory()]
ev = torch.tensor([1, 0, 1, 1]).bool().to(dev)
t = torch.tensor([5.0, 3.0, 8.0, 1.0], device=dev)
g = ghead.GraphedModelStep(c1, host[0].to(dev), host[1].to(dev), host[2].to(dev), None,
                           lambda hz, *_: gcox.neg_partial_log_likelihood(hz, ev, t, checks=False))
def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
def full():
    g.step(*host); opt.step()
print(f"PDL={os.environ.get('B200SURV_PDL', '1')}: step+opt {timed(full):.1f} us, step {timed(lambda: g.step(*host)):.1f} us, "
      f"replay {timed(g.replay):.1f} us, opt {timed(opt.step):.1f} us")
