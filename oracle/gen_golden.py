"""Generate tests/golden/*.npz by RUNNING the reference's own Python.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (``python -m oracle.gen_golden``): /root/reference does not exist
on the GPU box, so the vectors are committed.  Nothing is copied from the reference into this
repository: function and class definitions are located in the reference files by name with
``ast`` at run time, compiled in a scratch namespace, executed on seeded inputs, and only their
numeric inputs/outputs are stored.

What is pinned (reference file:line of the code that produced each vector):
* cox_fallback.npz   -- in-repo fallback losses
      partial_modality_training.py:296-311  (cox_loss, logcumsumexp form, /(sum(event)+1e-8))
      simple_fusion.py:47-57                (log(cumsum(exp)) form, /(sum(event)+1e-8))
      flexible_multimodal.py:43-52          (log(cumsum(exp)+1e-8))
      train_rnaseq_only.py:40-53            (/sum(event), no epsilon)
  on TIE-FREE times, where all of them equal the textbook no-ties Cox NLL averaged over events.
* cindex_fallback.npz -- fallback ConcordanceIndex, simple_fusion.py:59-73 (double loop).
* head_gated.npz / head_ungated.npz -- PartialModalityNet (partial_modality_training.py:165-277)
  and MultiModalSurvivalNet (final_multimodal.py:59-150) with rna_dim=40: full state_dict, inputs,
  eval-mode outputs, train-mode (dropout p forced to 0) outputs, parameter gradients of
  sum(hazard)+gate terms, BatchNorm running statistics after one train step.
* gate_entropy in head_gated.npz -- gate_entropy_loss, partial_modality_training.py:322-331.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/scripts/training"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _find(tree, name, kinds, pick=-1):
    hits = [n for n in ast.walk(tree) if isinstance(n, kinds) and n.name == name]
    hits.sort(key=lambda n: n.lineno)
    if not hits:
        raise KeyError(name)
    return hits[pick]


def extract(fname, name, kinds=(ast.FunctionDef, ast.ClassDef), pick=-1, extra_ns=None):
    """Compile one def/class out of a reference script without importing (= executing) the script."""
    path = os.path.join(REF, fname)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    node = _find(tree, name, kinds, pick)
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"torch": torch, "nn": nn, "np": np, "USE_MONAI": False}
    ns.update(extra_ns or {})
    exec(compile(mod, path, "exec"), ns)
    return ns[name], node.lineno


def tie_free_cohort(n, seed):
    g = torch.Generator().manual_seed(seed)
    time = torch.randperm(4000, generator=g)[:n].float() + 1.0        # distinct integer days
    event = (torch.rand(n, generator=g) < 0.4)
    if event.sum() == 0:
        event[0] = True
    log_hz = torch.randn(n, generator=g)
    return log_hz, event, time


def gen_cox():
    variants = {
        "partial_modality": ("partial_modality_training.py", "cox_loss", -1),
        "simple_fusion": ("simple_fusion.py", "neg_partial_log_likelihood", -1),
        "flexible": ("flexible_multimodal.py", "neg_partial_log_likelihood", -1),
        "rnaseq_only": ("train_rnaseq_only.py", "neg_partial_log_likelihood", -1),
    }
    fns = {}
    for k, (f, name, pick) in variants.items():
        fn, line = extract(f, name, (ast.FunctionDef,), pick)
        fns[k] = fn
        print(f"  cox variant {k}: {f}:{line}")
    out = {}
    cases = [("ka1", None)] + [(f"n{n}_s{s}", (n, s)) for n, s in
                               [(2, 1), (3, 2), (8, 3), (8, 4), (64, 5), (348, 6), (348, 7)]]
    names = []
    for cname, spec in cases:
        if spec is None:
            log_hz = torch.tensor([0.1, 0.5, -0.3, 0.2])
            event = torch.tensor([True, False, True, True])
            time = torch.tensor([5.0, 3.0, 8.0, 1.0])
        else:
            log_hz, event, time = tie_free_cohort(*spec)
        names.append(cname)
        out[f"{cname}/log_hz"] = log_hz.numpy()
        out[f"{cname}/event"] = event.numpy()
        out[f"{cname}/time"] = time.numpy()
        for k, fn in fns.items():
            for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
                x = log_hz.to(dt).clone().requires_grad_(True)
                loss = fn(x, event.to(dt), time.to(dt))
                loss.backward()
                out[f"{cname}/{k}/{tag}/loss"] = loss.detach().numpy()
                out[f"{cname}/{k}/{tag}/grad"] = x.grad.numpy()
    out["cases"] = np.array(names)
    out["variants"] = np.array(list(variants))
    np.savez_compressed(os.path.join(OUT, "cox_fallback.npz"), **out)


def gen_cindex():
    cls, line = extract("simple_fusion.py", "ConcordanceIndex", (ast.ClassDef,), -1)
    print(f"  cindex fallback: simple_fusion.py:{line}")
    out = {}
    names = []
    specs = [(22, 11, False, False), (116, 12, False, False), (116, 13, True, False),
             (116, 14, True, True), (348, 15, True, True), (5, 16, False, False)]
    for n, seed, time_ties, risk_ties in specs:
        g = torch.Generator().manual_seed(seed)
        if time_ties:
            time = torch.clamp(torch.floor(torch.empty(n).exponential_(1 / 30.0, generator=g)), 1, 100)
        else:
            time = torch.randperm(4000, generator=g)[:n].float() + 1
        event = torch.rand(n, generator=g) < 0.4
        est = torch.randn(n, generator=g)
        if risk_ties:
            est = torch.round(est * 4) / 4
        val = cls()(est, event.float(), time)
        cname = f"n{n}_s{seed}"
        names.append(cname)
        out[f"{cname}/est"] = est.numpy()
        out[f"{cname}/event"] = event.numpy()
        out[f"{cname}/time"] = time.numpy()
        out[f"{cname}/value"] = val.numpy()
    # fallback C-index 0.75
    est = torch.tensor([0.1, 0.5, -0.3, 0.2]); event = torch.tensor([1., 0., 1., 1.]); time = torch.tensor([5., 3., 8., 1.])
    names.append("ka1")
    out["ka1/est"], out["ka1/event"], out["ka1/time"] = est.numpy(), event.bool().numpy(), time.numpy()
    out["ka1/value"] = cls()(est, event, time).numpy()
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "cindex_fallback.npz"), **out)


def _set_dropout_p(model, p):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = p


def gen_head():
    rna_dim, B = 40, 6
    gated_cls, l1 = extract("partial_modality_training.py", "PartialModalityNet", (ast.ClassDef,))
    plain_cls, l2 = extract("final_multimodal.py", "MultiModalSurvivalNet", (ast.ClassDef,))
    gel, l3 = extract("partial_modality_training.py", "gate_entropy_loss", (ast.FunctionDef,))
    print(f"  heads: partial_modality_training.py:{l1}, final_multimodal.py:{l2}, gate_entropy:{l3}")
    for tag, cls, gated in (("gated", gated_cls, True), ("ungated", plain_cls, False)):
        torch.manual_seed(123 if gated else 321)
        model = cls(rna_dim=rna_dim, clinical_dim=1).double()
        # make BN parameters/statistics non-trivial so that eval mode exercises them
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, nn.BatchNorm1d):
                    m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.2, 0.2)
                    m.running_mean.uniform_(-0.3, 0.3); m.running_var.uniform_(0.5, 1.5)
        g = torch.Generator().manual_seed(7)
        ct = torch.rand(B, 1, 16, 16, 8, generator=g).double()
        rna = torch.randn(B, rna_dim, generator=g).double()
        clin = (0.3 + 0.6 * torch.rand(B, 1, generator=g)).double()
        mask = (torch.rand(B, 3, generator=g) < 0.7).double()
        mask[0] = 1.0
        out = {"rna": rna.numpy(), "clinical": clin.numpy(), "mask": mask.numpy(), "ct": ct.numpy()}
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        for k, v in sd0.items():
            if not k.startswith("ct_encoder"):      # the CT encoder is outside the head (contract a4)
                out["sd0/" + k] = v.numpy()
        args = (ct, rna, clin, mask) if gated else (ct, rna, clin)
        # the CT feature the head consumes (contract a4): ct_encoder(ct).view(B,-1)
        model.eval()
        with torch.no_grad():
            out["eval/ct_feat"] = model.ct_encoder(ct).view(B, -1).numpy()
            res = model(*args)
            if gated:
                out["eval/hazard"], out["eval/gate"] = res[0].numpy(), res[1].numpy()
                out["eval/gate_entropy"] = gel(res[1]).numpy()
            else:
                out["eval/hazard"] = res.numpy()
        # train mode, dropout disabled (Philox streams cannot be matched)
        model.train()
        _set_dropout_p(model, 0.0)
        feat = model.ct_encoder(ct).view(B, -1)
        out["train/ct_feat"] = feat.detach().numpy()
        # NB: calling ct_encoder twice would update BN3d stats twice; restore and run the real step
        model.load_state_dict(sd0)
        res = model(*args)
        if gated:
            hz, gate = res
            wsum = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
            obj = (hz * wsum).sum() + 0.01 * gel(gate)
            out["train/gate"] = gate.detach().numpy()
        else:
            hz = res
            wsum = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
            obj = (hz * wsum).sum()
        out["train/hazard"] = hz.detach().numpy()
        out["train/hazard_weights"] = wsum.numpy()
        obj.backward()
        for k, p_ in model.named_parameters():
            if not k.startswith("ct_encoder") and p_.grad is not None:
                out["grad/" + k] = p_.grad.numpy()
        for k, v in model.state_dict().items():
            if ("running" in k or "num_batches" in k) and not k.startswith("ct_encoder"):
                out["sd1/" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"head_{tag}.npz"), **out)
        # full-size key/shape manifest (a1): names and shapes only
        full = cls()
        man = {k: np.array(v.shape, dtype=np.int64) for k, v in full.state_dict().items()}
        np.savez_compressed(os.path.join(OUT, f"head_{tag}_manifest.npz"), **man)


def gen_ct_encoder():
    """The CT branch of the reference's PartialModalityNet (partial_modality_training.py:179-190, USE_MONAI=False) run
    through the reference class itself: eval features, one training step of the whole model (dropout off) and the
    gradients / running statistics of ct_encoder.*  (SURVEY.md 8f row 3)."""
    rna_dim, B = 40, 6
    cls, l1 = extract("partial_modality_training.py", "PartialModalityNet", (ast.ClassDef,))
    print(f"  ct encoder: partial_modality_training.py:{l1}")
    torch.manual_seed(123)
    model = cls(rna_dim=rna_dim, clinical_dim=1).double()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm3d)):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.2, 0.2)
                m.running_mean.uniform_(-0.3, 0.3); m.running_var.uniform_(0.5, 1.5)
    g = torch.Generator().manual_seed(11)
    ct = torch.rand(B, 1, 24, 20, 12, generator=g).double()
    ct[1].zero_()                                    # a patient without imaging (:89)
    rna = torch.randn(B, rna_dim, generator=g).double()
    clin = (0.3 + 0.6 * torch.rand(B, 1, generator=g)).double()
    mask = torch.ones(B, 3).double(); mask[1, 0] = 0.0
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    out = {"ct": ct.numpy().astype(np.float32), "rna": rna.numpy(), "clinical": clin.numpy(), "mask": mask.numpy()}
    for k, v in sd0.items():
        out["sd0/" + k] = v.numpy().astype(np.float32) if v.dtype.is_floating_point else v.numpy()
    model.eval()
    with torch.no_grad():
        out["eval/ct_feat"] = model.ct_encoder(ct).view(B, -1).numpy()
        out["eval/hazard"] = model(ct, rna, clin, mask)[0].numpy()
    model.train()
    _set_dropout_p(model, 0.0)
    hz, gate = model(ct, rna, clin, mask)
    wsum = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
    (hz * wsum).sum().backward()
    out["train/hazard"], out["train/hazard_weights"] = hz.detach().numpy(), wsum.numpy()
    for k, p_ in model.named_parameters():
        if k.startswith("ct_encoder"):
            out["grad/" + k] = p_.grad.numpy().astype(np.float32)
    for k, v in model.state_dict().items():
        if k.startswith("ct_encoder") and ("running" in k or "num_batches" in k):
            out["sd1/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "ct_encoder.npz"), **out)


def main():
    if "--only-ct" in sys.argv:
        gen_ct_encoder()
        return
    if not os.path.isdir(REF):
        sys.exit("reference not present: golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    print("generating golden vectors from", REF)
    gen_cox()
    gen_cindex()
    gen_head()
    gen_ct_encoder()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
