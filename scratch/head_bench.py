"""Head fwd+bwd at B=4096 (BASELINE.json configs[1]): wall/GPU time per step; `--ncu` keeps it short for a launch list."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import head as ghead, synth

dev = torch.device("cuda", 0)
hb = int(os.environ.get("B", 4096))
net = ghead.PartialModalityNet().to(dev).train()
hct, hrna, hclin, hmask = [x.to(dev) for x in synth.modality_batch(hb, seed=1234)]
hw = torch.randn(hb, device=dev) / hb ** 0.5


def head_step():
    for prm in net.parameters():
        prm.grad = None
    hz, gt = net.forward_features(hct, hrna, hclin, hmask)
    ((hz * hw).sum() + 0.01 * ghead.gate_entropy_loss(gt)).backward()


reps = 3 if "--ncu" in sys.argv else 30
for _ in range(3):
    head_step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(reps):
    head_step()
b.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"B={hb}: gpu {a.elapsed_time(b) / reps * 1e3:.1f} us/step, host enqueue {(t1 - t0) / reps * 1e6:.1f} us/step")
