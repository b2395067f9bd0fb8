import cProfile, pstats, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_survival_prediction_b200 import head as ghead, synth
dev = torch.device("cuda", 0)
net = ghead.PartialModalityNet().to(dev).train()
hct, hrna, hclin, hmask = [x.to(dev) for x in synth.modality_batch(4096, seed=1234)]
hw = torch.randn(4096, device=dev) / 64
def head_step():
    for prm in net.parameters():
        prm.grad = None
    hz, gt = net.forward_features(hct, hrna, hclin, hmask)
    ((hz * hw).sum() + 0.01 * ghead.gate_entropy_loss(gt)).backward()
for _ in range(5): head_step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): head_step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
