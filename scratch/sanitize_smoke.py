"""Small invocations of every kernel family in one script (a quick sanity run; compute-sanitizer is closed on this pool)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_survival_prediction_b200 as pkg
from multimodal_survival_prediction_b200 import synth, head as ghead
from multimodal_survival_prediction_b200.cindex import cindex_counts_shard, cindex_counts_cohorts
from multimodal_survival_prediction_b200.optim import ClipAdam

dev = torch.device("cuda", 0)
for n, mode in ((300, "small"), (20_000, "binned"), (9_000, "sorted")):
    lh, ev, t = synth.cohort(n, 7, few_ties=(mode == "sorted"))
    x = lh.to(dev).requires_grad_(True)
    loss = pkg.neg_partial_log_likelihood(x, ev.to(dev), t.to(dev), mode=mode)
    loss.backward()
    print(mode, float(loss.detach()))
lh, ev, t = synth.cohort(12_000, 9, risk_tie_frac=0.1)
print("cindex", float(pkg.ConcordanceIndex()(lh.to(dev), ev.to(dev), t.to(dev))))
print("shard", cindex_counts_shard(lh.to(dev), ev.to(dev), t.to(dev), 1, 3).tolist())
print("cohorts", cindex_counts_cohorts(lh.to(dev), ev.to(dev), t.to(dev), [0, 5000, 5001, 12_000]).sum().item())
off = torch.tensor([0, 5000, 9000, 12_000])
tt = torch.clamp(torch.floor(t / 10), 1, 4000)
xs = lh.to(dev).requires_grad_(True)
ls = pkg.neg_partial_log_likelihood_segmented(xs, ev.to(dev), tt.to(dev), off, mode="binned")
ls.sum().backward()
print("segmented", ls.tolist())
hs, ts, es, ne = pkg.select_labelled(lh.to(dev), torch.stack([t, ev.float()], 1).to(dev), (torch.arange(12_000) % 3 != 0))
print("compact", hs.shape[0], ne)
net = ghead.PartialModalityNet(rna_dim=40).to(dev).train()
ct, rna, clin, mask = [v.to(dev) for v in synth.modality_batch(64, rna_dim=40, seed=1)]
hz, gate = net.forward_features(ct, rna, clin, mask)
(hz.sum() + 0.01 * ghead.gate_entropy_loss(gate)).backward()
opt = ClipAdam([p for k, p in net.named_parameters() if p.grad is not None], lr=1e-3, weight_decay=1e-4)
opt.step()
torch.cuda.synchronize()
print("head + ClipAdam ok, total norm", float(opt.last_total_norm))
