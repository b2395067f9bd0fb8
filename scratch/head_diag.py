import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from multimodal_survival_prediction_b200 import head as ghead, synth
from oracle import head as ohead
import test_head_gpu as T

def rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return ((a - ref).norm() / (ref.norm() + 1e-30)).item(), ref.norm().item()

for gated in (True, False):
    for B, p_drop in ((6, 0.0), (300, 0.3), (4096, 0.3)):
        torch.manual_seed(B)
        m = (ghead.PartialModalityNet if gated else ghead.MultiModalSurvivalNet)().cuda()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = p_drop
        ct, rna, clin, mask = [t.cuda() for t in synth.modality_batch(B, seed=B)]
        if not gated: mask = None
        wts = torch.randn(B, generator=torch.Generator().manual_seed(B)).cuda() / B ** 0.5
        wts = wts - wts.mean()   # like a Cox gradient: sums to zero
        m.train()
        before = {k: v.clone() for k, v in m.state_dict().items()}
        ctg = ct.clone().requires_grad_(True)
        out = ghead.fused_head(m, ctg, rna, clin, mask, want_masks=True, seed=77)
        keep1, keep2 = out[-2].cpu(), out[-1].cpu()
        obj = (out[0] * wts).sum()
        if gated: obj = obj + 0.01 * ghead.gate_entropy_loss(out[1])
        obj.backward()
        m2 = (ghead.PartialModalityNet if gated else ghead.MultiModalSurvivalNet)(); m2.load_state_dict(before)
        hz_ref, gate_ref, p_ref, dct_ref, stats = T.oracle_run(m2, ct, rna, clin, mask, wts, True,
                                                              keep1 if p_drop > 0 else None, keep2 if p_drop > 0 else None, bf16=True)
        print(f"--- gated={gated} B={B} p={p_drop}: hazard rel {rel(out[0], hz_ref)[0]:.2e}", "gate rel %.2e" % rel(out[1], gate_ref)[0] if gated else "")
        print(f"    d_ct rel {rel(ctg.grad, dct_ref)[0]:.2e}")
        for k, v in m.named_parameters():
            if k.startswith("ct_encoder"): continue
            r, n = rel(v.grad, p_ref[k].grad)
            print(f"    {k:32s} rel {r:.2e}  |ref| {n:.2e}")
