"""One fwd+bwd of the CT encoder at B=64, 64x64x32 (for an ncu launch list)."""
import sys, torch
sys.path.insert(0, ".")
from multimodal_survival_prediction_b200.ctenc import CTEncoderCNN
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = CTEncoderCNN().to(dev).train()
ct = torch.rand(B, 1, 64, 64, 32, device=dev)
for _ in range(2):
    m(ct).sum().backward()
torch.cuda.synchronize()
