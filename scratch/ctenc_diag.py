"""Per-stage diagnostics of the CT encoder against torch (run on the GPU box): full-precision reference, a reference
with bf16-rounded convolution operands (what the tensor-core path computes), and timings against cuDNN."""
import copy, sys, torch
import torch.nn.functional as F
from torch import nn
sys.path.insert(0, ".")
from multimodal_survival_prediction_b200.ctenc import CTEncoderCNN
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
class RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.bfloat16().float()
    @staticmethod
    def backward(ctx, g): return g
def matched(seq, ct):
    """the reference CNN with the operands of convolutions 2 and 3 rounded to bf16 (straight-through)"""
    x = ct
    for i in (0, 3, 6):
        conv, bn = seq[i], seq[i + 1]
        if i == 0: x = conv(x)
        else: x = F.conv3d(RoundBf16.apply(x), RoundBf16.apply(conv.weight), conv.bias, stride=2, padding=1)
        x = F.relu(bn(x))
    return seq[9](x)
for shape, training in [((3, 1, 16, 16, 8), True), ((4, 1, 64, 64, 32), True), ((40, 1, 32, 32, 16), True)]:
    torch.manual_seed(0)
    ours = CTEncoderCNN().to(dev)
    ref = nn.Sequential(*[copy.deepcopy(m) for m in ours]).to(dev)
    ref2 = nn.Sequential(*[copy.deepcopy(m) for m in ours]).to(dev).double()
    refm = nn.Sequential(*[copy.deepcopy(m) for m in ours]).to(dev)
    for m in (ours, ref, ref2, refm): m.train(training)
    ct = torch.rand(shape, device=dev)
    w = torch.randn(shape[0], 128, 1, 1, 1, device=dev)
    yr = ref(ct); (yr * w).sum().backward()
    y2 = ref2(ct.double()); (y2 * w.double()).sum().backward()
    ym = matched(refm, ct); (ym * w).sum().backward()
    y = ours(ct); (y * w).sum().backward()
    torch.cuda.synchronize()
    print(shape, "out: ours vs fp32 %.2e, ours vs matched %.2e, fp32 vs fp64 %.2e" % (rel(y, yr), rel(y, ym), rel(yr, y2)))
    for (k, pa), (_, pb), (_, pc), (_, pd) in zip(ours.named_parameters(), ref.named_parameters(), ref2.named_parameters(), refm.named_parameters()):
        print("   grad %-9s ours vs fp32 %.4f  ours vs matched %.4f  matched vs fp32 %.4f  fp32 vs fp64 %.2e" % (k, rel(pa.grad, pb.grad), rel(pa.grad, pd.grad), rel(pd.grad, pb.grad), rel(pb.grad, pc.grad)))
for B in (4, 64):
    ours = CTEncoderCNN().to(dev).train(); ref = nn.Sequential(*[copy.deepcopy(m) for m in ours]).to(dev).train()
    ct = torch.rand(B, 1, 64, 64, 32, device=dev)
    refcl = copy.deepcopy(ref).to(memory_format=torch.channels_last_3d)
    def run_amp(m=refcl):
        with torch.autocast("cuda", dtype=torch.bfloat16): return m(ct)
    for name, fn, tf32 in (("ours", lambda: ours(ct), False), ("torch/cuDNN fp32", lambda: ref(ct), False), ("torch/cuDNN tf32", lambda: ref(ct), True), ("torch/cuDNN bf16 autocast channels_last", run_amp, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(3): fn().float().sum().backward()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn().float().sum().backward()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} {name}: fwd+bwd {e0.elapsed_time(e1)/10:.3f} ms")
