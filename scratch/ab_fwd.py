"""A/B of libb200surv builds in ONE process on ONE box: interleaved timing of the fused Cox forward (+ backward).

    B200SURV_PEER_TRACE=1 PYTHONPATH=. python scratch/ab_fwd.py scratch/ab/libA.so scratch/ab/libB.so ...

Box-to-box and thermal variation is larger than the differences of interest, so variants are timed round-robin
(ROUNDS x ITERS launches each) and the SM clock is sampled alongside."""
import ctypes
import os
import subprocess
import sys

import torch

from multimodal_survival_prediction_b200 import _lib as L
from multimodal_survival_prediction_b200 import synth

paths = sys.argv[1:] or [L.LIB_PATH]
n = int(os.environ.get("N", 1 << 24))
dev = torch.device("cuda", 0)
lh, ev, t = synth.cohort(n, 1234)
x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
grad = torch.empty(n, dtype=torch.float32, device=dev)
NAMES = ["start", "pass1", "sync1", "reduced", "sync2", "flag", "pulled", "lb1", "terms", "lb2", "end", "x11"]


def bind(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    for name, (res, args) in L.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


class Variant:
    def __init__(self, path):
        self.path, self.lib = path, bind(path)
        self.sb = self.lib.b200surv_cox_state_bytes(n, 1, L.COX_BINNED, 4096)
        self.wb = self.lib.b200surv_cox_workspace_bytes(n, 1, L.COX_BINNED, 4096)
        self.state = torch.zeros(self.sb, dtype=torch.uint8, device=dev)
        self.ws = torch.zeros(self.wb, dtype=torch.uint8, device=dev)
        self.loss = torch.empty(1, dtype=torch.float32, device=dev)
        self.one = torch.ones(1, dtype=torch.float32, device=dev)
        self.st = L.stream_ptr(dev)
        self.t_f, self.t_fb, self.t_b, self.seq = [], [], [], []

    def fwd(self):
        rc = self.lib.b200surv_cox_fwd(L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, 2, 0, L.COX_BINNED, 4096,
                                       ctypes.c_float(0.0), L.ptr(self.loss), L.ptr(self.state), self.sb,
                                       L.ptr(self.ws), self.wb, self.st)
        assert rc == 0, self.lib.b200surv_last_error()

    def bwd(self):
        rc = self.lib.b200surv_cox_bwd(L.ptr(self.one), L.ptr(self.state), self.sb, L.ptr(x), L.ptr(tt), L.ptr(e), None,
                                       n, 1, L.COX_BINNED, 4096, L.ptr(grad), self.st)
        assert rc == 0

    def trace(self):
        try:
            off = self.lib.b200surv_cox_peer_trace_offset(n, 4096)
        except Exception:
            return ""
        tr = self.ws[off:off + 8 * 32].view(torch.int64).cpu().tolist()
        out = ", ".join(f"{nm} {(v - tr[0]) / 1e3:.1f}" for nm, v in zip(NAMES, tr) if 0 <= v - tr[0] < 1e9)
        cy = tr[16:]
        if cy[0] and cy[3] > cy[0]:
            out += "\n   cycles (CTA 0 / block 0): " + ", ".join(f"{nm} {v - cy[0]}" for nm, v in zip(NAMES, cy) if 0 <= v - cy[0] < 1e9 and nm != "end")
            if tr[9] > tr[0]:
                out += f"   => SM clock {(cy[9] - cy[0]) / (tr[9] - tr[0]) * 1e3:.0f} MHz"
        if os.environ.get("BLOCK_TRACE"):
            bt = self.ws[off + 8 * 32:off + 8 * (32 + 8 * 128)].view(torch.int64).cpu().view(128, 8)
            rel = (bt[:, :4] - tr[0]).double() / 1e3
            out += "\n   per-block stamps (us): entry / A look-back done / G published / table written"
            for nm, fn in (("min", rel.min(0).values), ("median", rel.median(0).values), ("max", rel.max(0).values)):
                out += f"\n      {nm:6s} " + " ".join(f"{x:7.1f}" for x in fn.tolist())
            out += "\n      entry per block: " + " ".join(f"{x:.1f}" for x in rel[:, 0].tolist())
            out += "\n      lb1   per block: " + " ".join(f"{x:.1f}" for x in rel[:, 1].tolist())
            out += "\n      terms per block: " + " ".join(f"{x:.1f}" for x in (rel[:, 2] - rel[:, 1]).tolist())
            out += "\n      lb2   per block: " + " ".join(f"{x:.1f}" for x in (rel[:, 3] - rel[:, 2]).tolist())
        return out


def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception:
        return "?"


vs = [Variant(p) for p in paths]
for v in vs:
    for _ in range(5):
        v.fwd(); v.bwd()
torch.cuda.synchronize()
ROUNDS, ITERS = int(os.environ.get("ROUNDS", 6)), int(os.environ.get("ITERS", 20))
for r in range(ROUNDS):
    for v in vs:
        v.t_f.append(timed(v.fwd, ITERS))
        v.t_fb.append(timed(lambda: (v.fwd(), v.bwd()), ITERS))
        v.t_b.append(timed(v.bwd, ITERS))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * ITERS + 1)]     # per-kernel times inside the sequence
        evs[0].record()
        for i in range(ITERS):
            v.fwd(); evs[2 * i + 1].record(); v.bwd(); evs[2 * i + 2].record()
        torch.cuda.synchronize()
        fs = sorted(evs[2 * i].elapsed_time(evs[2 * i + 1]) * 1e3 for i in range(1, ITERS))
        bs = sorted(evs[2 * i + 1].elapsed_time(evs[2 * i + 2]) * 1e3 for i in range(1, ITERS))
        v.seq.append((fs[len(fs) // 2], bs[len(bs) // 2]))
    print(f"round {r}: clocks/power {clock()}  " + "  ".join(f"{os.path.basename(v.path)} f={v.t_f[-1]:.1f} fb={v.t_fb[-1]:.1f}" for v in vs), flush=True)
for v in vs:
    v.fwd(); torch.cuda.synchronize()
    f, fb = sorted(v.t_f), sorted(v.t_fb)
    b = sorted(v.t_b)
    sf, sb_ = sorted(x[0] for x in v.seq), sorted(x[1] for x in v.seq)
    print(f"{os.path.basename(v.path):24s} fwd median {f[len(f) // 2]:.1f} min {f[0]:.1f} us | fwd+bwd median {fb[len(fb) // 2]:.1f} min {fb[0]:.1f} us | "
          f"bwd alone {b[len(b) // 2]:.1f} | in sequence (events between calls): fwd {sf[len(sf) // 2]:.1f} bwd {sb_[len(sb_) // 2]:.1f} | loss {v.loss.item():.6f}")
    tr = v.trace()
    if tr and os.environ.get("B200SURV_PEER_TRACE"):
        print("   trace: " + tr)
