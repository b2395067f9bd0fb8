#!/usr/bin/env python
"""Headline benchmark: Cox NLL fwd+bwd patients/sec (BASELINE.json metric, configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--rows R]

Workload at N=1: Efron ties, 16,777,216 synthetic patients (integer days 1..4000 => heavy ties,
~30 % events, log_hz ~ N(0,1), seed 1234) resident in HBM; one step = forward + backward through the
C ABI (b200surv_cox_fwd: one cooperative launch; b200surv_cox_bwd: one launch).  At N>1 (torchrun, one rank per
GPU) each rank holds 16,777,216 rows of an N x 16M-row cohort: the per-bin int64 sums are exchanged INSIDE the forward
kernel over NVLink peer memory (b200surv_cox_binned_fwd_peer; --exchange nccl: partial / all-reduce / finalize)
each step (weak scaling; value = all rows / max-over-ranks time).

Extra measurements on the same JSON line: `e2e` (public Python API, pinned host inputs, H2D + D2H in
the timed region), `roofline` (dominant kernel, CUDA events), `cpu_baseline` (oracle port on the
host cores), `extra.cindex_1m` (C-index on 1M patients, tile-sharded over the N ranks), `extra.head_b4096` (gated
fusion head fwd+bwd, B = 4096), `extra.cv_sweep` (one GPU's share of the 5-fold CV sweep), `extra.ct_encoder` (the CNN CT
encoder fwd+bwd at B = 4 / 64 beside PyTorch/cuDNN on the same GPU), `extra.cfg1_batch4_step` (BASELINE.json configs[0]: one
training step of the ungated net at batch 4 from host inputs, as one CUDA graph and call by call, CPU port beside it).
`--impl reference` times the CPU port of the reference's loss (oracle/cox_torch.py) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = 1 << 24
SEED = 1234
ALGO_BYTES_PER_ROW = 22.0      # fwd 9 B read; bwd 9 B read + 4 B write
BWD_BYTES_PER_ROW = 13.0
FWD_BYTES_PER_ROW = 9.0
METRIC = "cox_nll_fwd_bwd_patients_per_sec"


def load_traffic(kernel, rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed summary of an
    `ncu --set full` capture of this command (profiles/ncu_traffic.json, written by scratch/ncu_traffic.py).  None when
    the capture was taken at another row count."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))[kernel]
        if int(d["rows"]) != int(rows):
            return None, None
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), d.get("source")
    except Exception:  # noqa: BLE001
        return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region, through NVML (a thread polling every ~0.5 ms: the timed
    region of the default run is a few milliseconds, far shorter than nvidia-smi's sampling period) plus one
    guaranteed sample right before and right after it; falls back to one `nvidia-smi` query when NVML is missing."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.bits, self.stop_flag, self.th, self.h, self.nv = gpu_index, [], 0, False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if i < len(ids) and ids[i].strip().isdigit():
                return int(ids[i])
        return i

    def sample(self):
        if self.nv is None:
            return
        try:
            self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            try:
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        except Exception:
            pass

    def _run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.0005)

    def start(self):
        self.sample()
        if self.nv is not None:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=2)
        self.sample()
        if self.nv is None or not self.samples:
            return self._smi_once()
        sm = sorted(self.samples)
        reasons = sorted(k for k, bit in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "source": "NVML, polled during the timed region"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": sorted(n for n, v in zip(names, f[2:]) if v.lower().startswith("active")), "samples": 1,
                    "source": "nvidia-smi, one query after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and may use the whole host."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_port_rate(rows, steps=1):
    """patients/s of the torch-CPU port (fp32, all host threads) on one pass over the workload."""
    import torch
    from multimodal_survival_prediction_b200 import synth
    from oracle import cox_torch
    use_all_host_threads()
    lh, ev, t = synth.cohort(rows, SEED)
    cox_torch.cox_nll_fwd_bwd(lh[: 1 << 16], ev[: 1 << 16], t[: 1 << 16])  # warm the thread pool
    t0 = time.perf_counter()
    for _ in range(steps):
        cox_torch.cox_nll_fwd_bwd(lh, ev, t)
    dt = (time.perf_counter() - t0) / steps
    return rows / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.rows            # the full workload: same config as the GPU arm (1.5-2 s per step on 16 threads)
    import torch
    from multimodal_survival_prediction_b200 import synth
    from oracle import cox_torch
    use_all_host_threads()
    lh, ev, t = synth.cohort(rows, SEED)
    for _ in range(max(args.warmup, 1)):
        cox_torch.cox_nll_fwd_bwd(lh[: 1 << 18], ev[: 1 << 18], t[: 1 << 18])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cox_torch.cox_nll_fwd_bwd(lh, ev, t)
    dt = (time.perf_counter() - t0) / args.steps
    val = rows / dt
    cores = torch.get_num_threads()
    sample = f"all {rows} rows of the workload per step, {args.steps} steps, torch CPU fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "patients/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cox_nll_fwd_bwd_efron_16M_heavy_ties", "rows_per_step": rows, "ties": "efron",
                   "note": "torchsurv is absent from the image; CPU port of the reference loss (oracle/cox_torch.py)"},
        "cpu_baseline": {"value": val, "unit": "patients/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "patients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def sorted_few_ties(n, dev, reps=5):
    """Cox Efron fwd+bwd on a FEW-TIES cohort (continuous times, SURVEY.md 8d variant): the SORTED path -- radix sort on
    time + look-back scans -- through b200surv_cox_fwd / _bwd with caller-owned buffers.  Returns (ms fwd, ms bwd, loss)."""
    import ctypes
    import torch
    from multimodal_survival_prediction_b200 import _lib as L
    from multimodal_survival_prediction_b200 import synth
    lib = L.load()
    lh, ev, t = synth.cohort(n, SEED, few_ties=True)
    x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
    sb = lib.b200surv_cox_state_bytes(n, 1, L.COX_SORTED, 0)
    wb = lib.b200surv_cox_workspace_bytes(n, 1, L.COX_SORTED, 0)
    state = torch.empty(sb, dtype=torch.uint8, device=dev)
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    one = torch.ones(1, dtype=torch.float32, device=dev)
    grad = torch.empty(n, dtype=torch.float32, device=dev)
    st = L.stream_ptr(dev)

    def fwd():
        L.check(lib.b200surv_cox_fwd(L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_SORTED, 0,
                                     ctypes.c_float(0.0), L.ptr(loss), L.ptr(state), sb, L.ptr(ws), wb, st), "b200surv_cox_fwd")

    def bwd():
        L.check(lib.b200surv_cox_bwd(L.ptr(one), L.ptr(state), sb, L.ptr(x), L.ptr(tt), L.ptr(e), None, n, 1, L.COX_SORTED, 0,
                                     L.ptr(grad), st), "b200surv_cox_bwd")

    for _ in range(3):
        fwd(); bwd()
    torch.cuda.synchronize()
    tf = tb = 0.0
    launches0 = lib.b200surv_debug_launch_count()
    for _ in range(reps):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); fwd(); b.record(); bwd(); c.record()
        torch.cuda.synchronize()
        tf += a.elapsed_time(b); tb += b.elapsed_time(c)
    sorted_few_ties.launches = int(lib.b200surv_debug_launch_count() - launches0)   # counted by the library at its launch sites
    return tf / reps, tb / reps, float(loss.item())


def run_few_ties(args):
    """--workload few_ties: the SORTED path as its own bench line.  N = 1: b200surv_cox_fwd / _bwd (mode SORTED).  N > 1 (torchrun):
    TIME-RANGE shards (dist.ShardedCoxSorted: three 128-byte records per rank all-gathered over NCCL, carry-in folded on the
    device); weak scaling (16,777,216 rows per rank, rank r's times in [r, r + 1) * T), a parity check of the sharded result
    against one GPU on the same cohort, and the strong-scaling figure."""
    import torch
    import torch.distributed as dist
    from multimodal_survival_prediction_b200 import _lib as L
    from multimodal_survival_prediction_b200 import cox as gcox
    from multimodal_survival_prediction_b200 import dist as gdist
    from multimodal_survival_prediction_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.rows
    steps = max(args.steps, 3)
    sampler = ClockSampler(local_rank)
    parity = strong = None
    if world == 1:
        sampler.start()
        fwd_ms, bwd_ms, loss = sorted_few_ties(n, dev, reps=steps)
        clocks = sampler.stop()
        ms = fwd_ms + bwd_ms
        launches = getattr(sorted_few_ties, "launches", None)
    else:
        def barrier():
            dist.barrier()
            torch.cuda.synchronize()

        def timed(op, x, tt, e, grad):
            for _ in range(3):
                op.forward(x, tt, e); op.backward(grad)
            barrier()
            l0 = L.load().b200surv_debug_launch_count()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            for _ in range(steps):
                op.forward(x, tt, e); op.backward(grad)
            e1.record()
            nl = int(L.load().b200surv_debug_launch_count() - l0)
            barrier()
            tm = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tm.item()), nl

        # weak scaling: every rank its own 16.7M rows; the global cohort's time axis is cut into one range per rank
        lh, ev, t = synth.cohort(n, SEED + rank, few_ties=True)
        t = t + rank * 20000.0                       # Exp(1000) never reaches 20000 at this n: rank r's times lie in its own range
        x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
        grad = torch.empty(n, dtype=torch.float32, device=dev)
        op = gdist.ShardedCoxSorted(n, dev, ties="efron")
        sampler.start()
        ms, launches = timed(op, x, tt, e, grad)
        clocks = sampler.stop()
        if op.check():
            raise SystemExit("few_ties: flags raised on synthetic data")
        loss = float(op.loss.item())
        fwd_ms = bwd_ms = None
        del op, x, e, tt, grad
        # parity + strong scaling: the SAME 16.7M-row cohort, sorted by time on the host and cut into contiguous time ranges
        glh, gev, gt = synth.cohort(n, SEED, few_ties=True)
        order = torch.argsort(gt, stable=True)
        a_, b_ = gdist.shard_bounds(n, rank, world)
        idx = order[a_:b_][torch.randperm(b_ - a_, generator=torch.Generator().manual_seed(rank))]   # a shard is an unordered row set
        sx, se, st_ = glh[idx].to(dev), gev[idx].to(dev), gt[idx].to(dev)
        sgrad = torch.empty(b_ - a_, dtype=torch.float32, device=dev)
        sop = gdist.ShardedCoxSorted(b_ - a_, dev, ties="efron")
        sms, _ = timed(sop, sx, st_, se, sgrad)
        ref = torch.zeros(1, dtype=torch.float32, device=dev)
        if rank == 0:
            fx, fe, ft = glh.to(dev), gev.to(dev), gt.to(dev)
            l1, state1 = gcox.cox_fwd_raw(fx, ft, fe, None, 1, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_SORTED, 0)
            full_grad = gcox.cox_bwd_raw(torch.ones(1, device=dev), state1, fx, ft, fe, None, 1, L.COX_SORTED, 0)
            ref[0] = l1[0]
        dist.broadcast(ref, 0)
        loss_rel = abs(float(sop.loss.item()) - float(ref.item())) / max(abs(float(ref.item())), 1e-30)
        didx = idx.to(dev)
        if rank == 0:
            gmax = float(full_grad.abs().max())
            err = float((sgrad - full_grad[didx]).abs().max())
            for r in range(1, world):
                ra, rb = gdist.shard_bounds(n, r, world)
                gi = torch.empty(rb - ra, dtype=torch.int64, device=dev); gg = torch.empty(rb - ra, dtype=torch.float32, device=dev)
                dist.recv(gi, src=r); dist.recv(gg, src=r)
                err = max(err, float((gg - full_grad[gi]).abs().max()))
            grad_rel = err / gmax
        else:
            dist.send(didx, dst=0); dist.send(sgrad, dst=0)
            grad_rel = 0.0
        ok = torch.tensor([int(loss_rel <= 2e-6 and grad_rel <= 2e-6)], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity = {"ok": bool(ok.item()), "loss_rel_diff": loss_rel, "grad_max_abs_diff_over_max": grad_rel, "tolerance": 2e-6,
                  "rows_total": n, "exchange": "3 x all_gather of one 128-byte record per rank (NCCL)",
                  "what": "the same 16,777,216-row few-ties cohort cut into time-range shards over the ranks vs one GPU (rank 0, mode sorted); "
                          "fp64 sums in another order, so a tolerance and not bit equality"}
        strong = {"rows_total": n, "rows_per_gpu": b_ - a_, "ms_per_step": sms, "value": n / (sms * 1e-3), "unit": "patients/s",
                  "scaling": "strong"}
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_check": parity, "error": "sharded result differs from one GPU"}))
            dist.destroy_process_group()
            raise SystemExit(3)
    peak, peak_src = load_peaks()
    achieved = ALGO_BYTES_PER_ROW * n / (ms * 1e-3) / 1e9          # per GPU
    traffic, traffic_src = load_traffic("cox_sorted_step", n)     # all kernels of one fwd+bwd step, ncu --set full
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": world * n / (ms * 1e-3), "unit": "patients/s", "n_gpus": world, "steps": steps, "warmup": 3,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums over f32 inputs",
            "data": "synthetic",
            "config": {"workload": "cox_nll_fwd_bwd_efron_16M_few_ties", "rows_per_gpu": n, "ties": "efron", "time": "Exp(1000), continuous",
                       "event_rate": 0.30, "mode": "sorted", "l2": "inputs and scratch (68 B/row) far larger than L2",
                       "parallelism": "single GPU" if world == 1 else f"{world} time-range shards, boundary records all-gathered (NCCL)"},
            "loss": loss,
            "roofline": {"bound": "hbm", "scope": "step = sort + scans + gradient scatter, 22 algorithmic B/row, per GPU", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "note": "the path is bound by instruction issue and the LSU pipe (radix ranking, fp64 struct scans), not by DRAM: "
                                 "traffic / ms is a fraction of the HBM peak (DESIGN.md 3.2)"},
            "gpu_launches": launches, "parity_check": parity, "strong_scaling": strong, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def run_b200(args):
    import torch
    import torch.distributed as dist
    from multimodal_survival_prediction_b200 import _lib as L
    from multimodal_survival_prediction_b200 import cindex as gci
    from multimodal_survival_prediction_b200 import cox as gcox
    from multimodal_survival_prediction_b200 import dist as gdist
    from multimodal_survival_prediction_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.require_device(local_rank)

    n = args.rows
    lh, ev, t = synth.cohort(n, SEED + rank)       # each rank: its own 16M-row block of the global cohort
    pin = [x.pin_memory() for x in (lh, ev, t)]
    x, e, tt = lh.to(dev), ev.to(dev), t.to(dev)
    grad = torch.empty(n, dtype=torch.float32, device=dev)
    op = gdist.ShardedCoxBinned(n, dev, nbins=4096, ties="efron", exchange=args.exchange)
    fused = world == 1 or op.exchange == "peer"   # forward is one cooperative launch

    def step():
        op.forward(x, tt, e)
        op.backward(x, tt, e, grad)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    hdr = gcox.read_headers(op.state, 1, L.COX_BINNED)[0]
    if hdr.flags != 0:
        raise SystemExit(f"binned precondition violated on synthetic data: flags={hdr.flags}")

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.load().b200surv_debug_launch_count()
    ev0.record()
    for i in range(args.steps):      # the headline loop: exactly K steps, nothing else on the stream
        op.forward(x, tt, e)
        op.backward(x, tt, e, grad)
    ev1.record()
    gpu_launches = int(L.load().b200surv_debug_launch_count() - launches0)   # counted by the library at its launch sites
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ms_step = ms_total / args.steps
    # per-kernel durations for the roofline: same loop again with CUDA events around each call
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
             torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        a, b, c = k_ev[i]
        a.record()
        op.forward(x, tt, e)
        b.record()
        op.backward(x, tt, e, grad)
        c.record()
    barrier()
    clocks = sampler.stop()
    fwd_ms = sum(a.elapsed_time(b) for a, b, _ in k_ev) / args.steps
    bwd_ms = sum(b.elapsed_time(c) for _, b, c in k_ev) / args.steps
    tmax = torch.tensor([ms_step], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item())
    value = world * n / (ms_step * 1e-3)
    loss_val = float(op.loss.item())

    # ---- N > 1: (a) PARITY of the sharded path, where the driver can see it: all ranks shard the SAME 16,777,216-row cohort
    # (contiguous row blocks), rank 0 also computes it on one GPU, every rank compares the loss bit for bit and the gradient
    # of its rows element for element; (b) STRONG scaling: the same total rows over N ranks, timed like the headline loop.
    parity, strong = None, None
    if world > 1:
        glh, gev, gt = synth.cohort(n, SEED)
        a_, b_ = gdist.shard_bounds(n, rank, world)
        sx, se, st_ = glh[a_:b_].to(dev), gev[a_:b_].to(dev), gt[a_:b_].to(dev)
        sgrad = torch.empty(b_ - a_, dtype=torch.float32, device=dev)
        sop = gdist.ShardedCoxBinned(b_ - a_, dev, nbins=4096, ties="efron", exchange=args.exchange)
        for _ in range(3):
            sop.forward(sx, st_, se)
            sop.backward(sx, st_, se, sgrad)
        barrier()
        ref = torch.zeros(2, dtype=torch.float32, device=dev)          # rank 0: loss of the whole cohort on ONE GPU
        full_grad = None
        if rank == 0:
            fx, fe, ft = glh.to(dev), gev.to(dev), gt.to(dev)
            l1, state1 = gcox.cox_fwd_raw(fx, ft, fe, None, 1, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_BINNED, 4096)
            full_grad = gcox.cox_bwd_raw(torch.ones(1, device=dev), state1, fx, ft, fe, None, 1, L.COX_BINNED, 4096)
            ref[0] = l1[0]
        dist.broadcast(ref, 0)
        loss_equal = bool(torch.equal(sop.loss.view(torch.int32), ref[:1].view(torch.int32)))
        # gradients: every rank sends its shard to rank 0 (NCCL), which compares them with the single-GPU gradient
        shards = [torch.empty(gdist.shard_bounds(n, r, world)[1] - gdist.shard_bounds(n, r, world)[0],
                              dtype=torch.float32, device=dev) for r in range(world)] if rank == 0 else None
        if rank == 0:
            shards[0].copy_(sgrad)
            for r in range(1, world):
                dist.recv(shards[r], src=r)
            grad_equal = bool(torch.equal(torch.cat(shards), full_grad))
            max_abs = float((torch.cat(shards) - full_grad).abs().max())
            del shards, full_grad, fx, fe, ft, state1
        else:
            dist.send(sgrad, dst=0)
            grad_equal, max_abs = True, 0.0
        ok = torch.tensor([int(loss_equal), int(grad_equal)], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity = {"loss_bit_equal": bool(ok[0].item()), "grad_equal": bool(ok[1].item()), "grad_max_abs_diff": max_abs,
                  "rows_total": n, "exchange": sop.exchange,
                  "what": "the same 16,777,216-row cohort row-sharded over the ranks vs one GPU (rank 0), fp32 loss bits and every gradient element"}
        # strong scaling, timed on the device, max over ranks
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for _ in range(args.steps):
            sop.forward(sx, st_, se)
            sop.backward(sx, st_, se, sgrad)
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1) / args.steps], device=dev)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"rows_total": n, "rows_per_gpu": b_ - a_, "ms_per_step": float(sms.item()),
                  "value": n / (float(sms.item()) * 1e-3), "unit": "patients/s", "scaling": "strong"}
        if sop.peers is not None:
            barrier()
            sop.peers.close()
        del sop, sx, se, st_, sgrad, glh, gev, gt
        if not (parity["loss_bit_equal"] and parity["grad_equal"]):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_check": parity, "error": "sharded result differs from one GPU"}))
            dist.destroy_process_group()
            raise SystemExit(3)

    if args.skip_extras:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "patients/s", "n_gpus": world,
                              "steps": args.steps, "ms_per_step": ms_step, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                              "exchange": op.exchange if world > 1 else None, "loss": loss_val,
                              "parity_check": parity, "strong_scaling": strong, "gpu_launches": gpu_launches,
                              "note": "--skip-extras: partial line, not a bench result"}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: public Python API, pinned host inputs -> H2D -> fwd (auto mode) -> bwd -> loss D2H
    def e2e_step():
        xd = pin[0].to(dev, non_blocking=True).requires_grad_(True)
        ed = pin[1].to(dev, non_blocking=True)
        td = pin[2].to(dev, non_blocking=True)
        loss = gcox.neg_partial_log_likelihood(xd, ed, td)
        loss.backward()
        return float(loss.item()), xd.grad

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lv, _g = e2e_step()
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_dt], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * n / float(e2e_t.item())

    # ---- C-index, 1M patients, row-sharded over the ranks (strong scaling), bit-exact int64 counts
    cn = 1 << 20
    clh, cev, ct = synth.cohort(cn, SEED)
    cx, ce, ctt = clh.to(dev), cev.to(dev), ct.to(dev)
    def timed_cindex(algo):
        for _ in range(2):
            gdist.cindex_counts_sharded(cx, ce, ctt, algo=algo)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        c0.record()
        for _ in range(reps):
            cnt = gdist.cindex_counts_sharded(cx, ce, ctt, algo=algo)
        c1.record()
        barrier()
        ms_ = torch.tensor([c0.elapsed_time(c1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return float(ms_.item()), cnt.cpu().tolist()

    # algo 1: the tiled pair-by-pair kernel BASELINE.json's north_star describes (and its 8-GPU scaling target is quoted on);
    # algo 2 (the library's default): ranks in sorted tiles, the same six integers
    ci_ms, counts = timed_cindex(1)
    ci2_ms, counts2 = timed_cindex(2)
    if counts2 != counts:
        raise SystemExit(f"C-index: algo 2 counters {counts2} differ from algo 1 {counts}")

    # ---- gated fusion head fwd+bwd, B = 4096 rows, bf16 tcgen05 GEMMs (BASELINE.json configs[1]); per rank
    from multimodal_survival_prediction_b200 import head as ghead
    hb = 4096
    net = ghead.PartialModalityNet().to(dev).train()
    hct, hrna, hclin, hmask = [x.to(dev) for x in synth.modality_batch(hb, seed=SEED)]
    hw = torch.randn(hb, device=dev) / hb ** 0.5

    def head_step():
        for prm in net.parameters():
            prm.grad = None
        hz, gt = net.forward_features(hct, hrna, hclin, hmask)
        (torch.dot(hz, hw) + 0.01 * ghead.gate_entropy_loss(gt)).backward()

    def timed_head(fn, reps=10):
        for _ in range(3):
            fn()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(reps):
            fn()
        h1.record()
        barrier()
        return h0.elapsed_time(h1) / reps

    head_eager_ms = timed_head(head_step)
    # the same step as ONE CUDA graph (head.GraphedHeadStep): the eager step is bound by the host (~100 launches)
    graphed = ghead.GraphedHeadStep(net, hct, hrna, hclin, hmask,
                                    lambda hz, gt: torch.dot(hz, hw) + 0.01 * ghead.gate_entropy_loss(gt))
    fresh = [x.clone() for x in (hct, hrna, hclin, hmask)]       # a "new batch": copied into the static buffers each step
    head_ms = timed_head(lambda: graphed.step(*fresh), reps=20)
    head_flop_per_row = 11_396_224          # fwd 5,507,136 + bwd 5,889,088

    # ---- CT encoder feeding the head: the reference's 3-conv CNN on 64 x 64 x 32 volumes, fwd + bwd,
    # at the reference's batch size (4, partial_modality_training.py:366) and at 64; PyTorch/cuDNN (bf16 autocast,
    # channels_last_3d, TF32 allowed: its fastest setting here) on the same GPU beside it
    import copy as _copy
    ct_extra = {}
    for cb in (4, 64):
        enc = net.ct_encoder.train()
        vol = torch.rand(cb, 1, 64, 64, 32, device=dev)
        cudnn_enc = torch.nn.Sequential(*[_copy.deepcopy(m_) for m_ in enc]).to(dev).to(memory_format=torch.channels_last_3d).train()

        def ct_step(enc=enc, vol=vol):
            for prm in enc.parameters():
                prm.grad = None
            enc(vol).sum().backward()

        def cudnn_step(m_=cudnn_enc, vol=vol):
            for prm in m_.parameters():
                prm.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y_ = m_(vol)
            y_.float().sum().backward()

        ct_extra[f"b{cb}"] = {"ms_fwd_bwd": timed_head(ct_step), "cudnn_bf16_channels_last_ms": timed_head(cudnn_step)}
    # ---- BASELINE.json configs[0]: one training step of the ungated net at the reference's batch size (4): CT volume +
    # 5,005 genes + age from pinned host memory -> CT encoder -> head -> Cox loss -> backward -> clip + AdamW, all through
    # the public modules; beside it the same step as a torch-CPU port on the host cores (rank 0, N=1 only)
    from multimodal_survival_prediction_b200.optim import ClipAdam
    c1 = ghead.MultiModalSurvivalNet().to(dev).train()
    c1_opt = ClipAdam(c1.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=1.0, adamw=True)      # simple_fusion.py:391
    _, c1_rna, c1_clin, _ = synth.modality_batch(4, seed=SEED)
    c1_host = [torch.rand(4, 1, 64, 64, 32).pin_memory(), c1_rna.pin_memory(), c1_clin.pin_memory()]
    c1_ev = torch.tensor([1, 0, 1, 1]).bool().to(dev)
    c1_t = torch.tensor([5.0, 3.0, 8.0, 1.0], device=dev)

    def cfg1_step():
        vol, rna_, clin_ = (x.to(dev, non_blocking=True) for x in c1_host)
        c1_opt.zero_grad()
        loss_ = gcox.neg_partial_log_likelihood(c1(vol, rna_, clin_), c1_ev, c1_t)
        loss_.backward()
        c1_opt.step()
        return loss_

    cfg1_eager_ms = timed_head(cfg1_step, reps=20)
    # the same step with forward + loss + backward replayed as ONE CUDA graph (head.GraphedModelStep); the batch is still
    # copied from pinned host memory and the optimizer still steps every iteration
    c1_graph = ghead.GraphedModelStep(c1, c1_host[0].to(dev), c1_host[1].to(dev), c1_host[2].to(dev), None,
                                      lambda hz, *_: gcox.neg_partial_log_likelihood(hz, c1_ev, c1_t, checks=False))

    def cfg1_graphed_step():
        loss_, _ = c1_graph.step(*c1_host)
        c1_opt.step()
        return loss_

    cfg1_ms = timed_head(cfg1_graphed_step, reps=20)
    cfg1 = {"ms_per_step": cfg1_ms, "ms_per_step_eager": cfg1_eager_ms, "patients_per_s": 4 / (cfg1_ms * 1e-3), "h2d_bytes_per_step": sum(x.numel() * 4 for x in c1_host),
            "note": "MultiModalSurvivalNet (CNN CT encoder + ungated head) + Cox loss + backward + fused clip/AdamW at batch 4, "
                    "64x64x32 CT + 5005 genes + age copied from pinned host memory every step; forward + loss + backward replayed as one "
                    "CUDA graph (head.GraphedModelStep), ms_per_step_eager = call by call (host-bound); per rank"}
    if rank == 0 and world == 1:
        try:
            from oracle import model_torch
            cm = model_torch.MultiModalNetCPU().train()
            copt = torch.optim.AdamW(cm.parameters(), lr=1e-4, weight_decay=1e-4)
            cargs = [x.clone() for x in c1_host] + [c1_ev.cpu(), c1_t.cpu()]
            for _ in range(2):
                model_torch.training_step(cm, copt, *cargs)
            t0 = time.perf_counter()
            for _ in range(10):
                model_torch.training_step(cm, copt, *cargs)
            cfg1["cpu_port_ms_per_step"] = (time.perf_counter() - t0) / 10 * 1e3
            cfg1["cpu_port"] = f"torch CPU fp32 port of the same step (oracle/model_torch.py), {torch.get_num_threads()} threads, 10 steps"
        except Exception as ex:  # noqa: BLE001 -- the baseline is optional
            cfg1["cpu_port"] = "unavailable: " + repr(ex)
    ct_extra["note"] = ("Conv3d(1,32)/(32,64)/(64,128) k3 s2 p1 + BatchNorm3d + ReLU + AdaptiveAvgPool3d(1) on 64x64x32 volumes, "
                        "training mode: direct first conv, im2col + tcgen05 GEMM for the other two (b200surv_ct_encoder_fwd/_bwd); per rank")

    # ---- CV sweep share of one GPU (BASELINE.json configs[4]): 32 replicas x one fold of 100k patients, packed back to
    # back: segmented Cox fwd+bwd (one call) + one C-index per replica.  Replicas are independent: no collective.
    sw_rep, sw_rows = 32, 100_000
    slh, sev, st_ = synth.cohort(sw_rep * sw_rows, 1000 + rank)
    sx, se_, stt = slh.to(dev), sev.to(dev), st_.to(dev)
    soff = torch.arange(0, sw_rep + 1, dtype=torch.int64) * sw_rows
    soff_d = soff.to(dev)
    sones = torch.ones(sw_rep, dtype=torch.float32, device=dev)

    def sweep_cox():
        loss, state = gcox.cox_fwd_raw(sx, stt, se_, soff_d, sw_rep, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_BINNED, 4096)
        return loss, gcox.cox_bwd_raw(sones, state, sx, stt, se_, soff_d, sw_rep, L.COX_BINNED, 4096)

    def sweep_ci():
        return gci.cindex_counts_cohorts(sx, se_, stt, soff)

    def timed_ms(fn, reps=5):
        for _ in range(2):
            fn()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        barrier()
        tt_ = torch.tensor([a0.elapsed_time(a1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        return float(tt_.item())

    sweep_cox_ms, sweep_ci_ms = timed_ms(sweep_cox), timed_ms(sweep_ci)

    # ---- BASELINE.json configs[4] END TO END: 256 fusion-head replicas x 5 folds x 100k patients across 8 GPUs = 32 replicas
    # per GPU (with N ranks: 32 x N replicas, i.e. the full sweep at N = 8).  Per fold the rank stages the fold's RNA matrix
    # once (bf16, shared by its replicas), runs every replica's gated head forward on the fold's 100k rows (eval mode), then
    # ONE segmented Cox fwd+bwd over the 32 packed hazard vectors and ONE C-index call over the 32 packed cohorts; at the end
    # one all-gather of the replica-fold C-indices.  Replicas are independent: no other collective.
    folds = 5
    fold_nets = [ghead.PartialModalityNet().to(dev).eval() for _ in range(sw_rep)]
    # the replicas' forward passes run on SWEEP_STREAMS streams (one staged copy of the fold's RNA matrix per stream): one
    # replica's tensor-core GEMM overlaps the others' memory-bound layers
    SWEEP_STREAMS = 4
    sweep_streams = [torch.cuda.Stream(device=dev) for _ in range(SWEEP_STREAMS)]
    fold_saved = [ghead.head_saved_buffer(sw_rows, 5005, dev) for _ in range(SWEEP_STREAMS)]
    g_ = torch.Generator(device=dev).manual_seed(4321 + rank)
    fold_x = torch.randn(sw_rows, 5005, device=dev, generator=g_)
    fold_ct = torch.relu(torch.randn(sw_rows, 128, device=dev, generator=g_))
    fold_clin = 0.3 + 0.6 * torch.rand(sw_rows, 1, device=dev, generator=g_)
    fold_mask = (torch.rand(sw_rows, 3, device=dev, generator=g_) < torch.tensor([142 / 608, 427 / 608, 587 / 608], device=dev)).float()
    fold_labels = []
    for f in range(folds):
        _, fev, ft = synth.cohort(sw_rows, 1000 + f)
        fold_labels.append((fev.to(dev).repeat(sw_rep), ft.to(dev).repeat(sw_rep)))
    hz_packed = torch.empty(sw_rep * sw_rows, dtype=torch.float32, device=dev)

    def full_sweep():
        counts = []
        losses = []
        with torch.no_grad():
            for f in range(folds):
                # (synthetic: the same feature matrices stand for every fold; the labels differ per fold)
                main_ = torch.cuda.current_stream(dev)
                for st_, sv_ in zip(sweep_streams, fold_saved):
                    st_.wait_stream(main_)                 # (the previous fold's loss / C-index read hz_packed)
                    with torch.cuda.stream(st_):
                        ghead.stage_rna(fold_x, sv_)
                for r_, net_ in enumerate(fold_nets):
                    with torch.cuda.stream(sweep_streams[r_ % SWEEP_STREAMS]):
                        hz_, _ = ghead.fused_head(net_, fold_ct, fold_x, fold_clin, fold_mask,
                                                  staged_saved=fold_saved[r_ % SWEEP_STREAMS])
                        hz_packed[r_ * sw_rows:(r_ + 1) * sw_rows].copy_(hz_)
                for st_ in sweep_streams:
                    main_.wait_stream(st_)
                fev_, ft_ = fold_labels[f]
                loss_, state_ = gcox.cox_fwd_raw(hz_packed, ft_, fev_, soff_d, sw_rep, L.TIES["efron"], L.REDUCE_MEAN_TERMS, L.COX_BINNED, 4096)
                gcox.cox_bwd_raw(sones, state_, hz_packed, ft_, fev_, soff_d, sw_rep, L.COX_BINNED, 4096)
                losses.append(loss_)
                counts.append(gci.cindex_counts_cohorts(hz_packed, fev_, ft_, soff))
        return torch.stack(losses), torch.stack(counts)      # [folds][replicas], [folds][replicas][6]

    full_sweep()
    barrier()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    w0.record()
    sweep_losses, sweep_counts = full_sweep()
    w1.record()
    barrier()
    sweep_wall = time.perf_counter() - t_wall0
    sweep_ms = torch.tensor([w0.elapsed_time(w1)], device=dev)
    if world > 1:
        dist.all_reduce(sweep_ms, op=dist.ReduceOp.MAX)
    sweep_ms = float(sweep_ms.item())
    ci_local = torch.tensor([[gci.cindex_from_counts(c) for c in fold_c] for fold_c in sweep_counts.cpu().tolist()], dtype=torch.float64,
                            device=dev)                       # [folds][replicas of this rank]
    if world > 1:
        gathered = [torch.empty_like(ci_local) for _ in range(world)]
        dist.all_gather(gathered, ci_local)                   # the one collective of the sweep: N x 160 C-indices
        ci_all = torch.cat(gathered, dim=1)
    else:
        ci_all = ci_local
    sweep_full = {"replicas": sw_rep * world, "folds": folds, "rows_per_fold": sw_rows, "replica_folds": int(ci_all.numel()),
                  "ms_gpu_max_over_ranks": sweep_ms, "wall_s_rank0": sweep_wall,
                  "replica_folds_per_s": ci_all.numel() / (sweep_ms * 1e-3),
                  "c_index_mean": float(ci_all.mean()), "c_index_min": float(ci_all.min()), "c_index_max": float(ci_all.max()),
                  "loss_mean": float(sweep_losses.mean()),
                  "note": "per replica-fold: gated head forward on 100k rows (eval; the fold's RNA matrix staged once in bf16), then per "
                          "fold one segmented Cox fwd+bwd over the packed hazards and one packed C-index call; all-gather of the "
                          "C-indices at the end; 32 replicas per GPU (256 at 8 GPUs = BASELINE.json configs[4])"}
    del fold_nets, fold_saved, fold_x, fold_ct, fold_clin, fold_mask, fold_labels, hz_packed

    # ---- the SORTED path on the few-ties variant of the headline cohort (continuous times), this rank's GPU
    ft_fwd_ms, ft_bwd_ms, ft_loss = sorted_few_ties(n, dev, reps=3)

    # ---- CPU baseline of the C-index (rank 0, N=1 only; bounded samples): the oracle's C brute force -- the same
    # O(n^2) pair rule as the GPU kernel and the reference's fallback (simple_fusion.py:59-73), all host cores (OpenMP)
    ci_cpu = None
    if rank == 0 and world == 1:
        try:
            from oracle import cindex as oci
            m = 40_000
            t0 = time.perf_counter()
            cb = oci.counts_brute(clh[:m].numpy(), cev[:m].numpy(), ct[:m].numpy())
            dt_b = time.perf_counter() - t0
            ci_cpu = {"kind": "port", "cores": os.cpu_count(),
                      "sample": f"first {m} of the 1,048,576 patients, brute-force C with OpenMP (oracle/cindex_oracle.c)",
                      "ordered_pairs_per_s": float(sum(int(v) for v in cb)) / dt_b, "seconds": dt_b,
                      "extrapolated_ms_at_1m": dt_b * (cn / m) ** 2 * 1e3}
        except Exception as ex:  # noqa: BLE001 -- the baseline is optional, the bench line is not
            ci_cpu = {"unavailable": repr(ex)}

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved_bwd = BWD_BYTES_PER_ROW * n / (bwd_ms * 1e-3) / 1e9
        achieved_fwd = FWD_BYTES_PER_ROW * n / (fwd_ms * 1e-3) / 1e9
        achieved_step = ALGO_BYTES_PER_ROW * n / (ms_step * 1e-3) / 1e9
        cpu_rows = n
        cpu_val, cpu_dt, cores = cpu_port_rate(cpu_rows) if world == 1 else (None, None, None)
        fwd_traffic, traffic_src = load_traffic("cox_binned_fwd_fused", n) if fused else (None, None)
        bwd_traffic, _ = load_traffic("cox_binned_bwd", n)
        out = {
            "metric": METRIC, "value": value, "unit": "patients/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cox_nll_fwd_bwd_efron_16M_heavy_ties", "rows_per_gpu": n, "ties": "efron",
                       "time": "integer days clamp(floor(Exp(1000)),1,4000)", "event_rate": 0.30, "mode": "binned",
                       "nbins": 4096,
                       "parallelism": "single GPU" if world == 1 else
                       (f"row-block shards x{world}, per-bin int64 sums exchanged inside the forward kernel over NVLink peer memory"
                        if op.exchange == "peer" else f"row-block shards x{world}, NCCL all-reduce of per-bin int64 sums"),
                       "l2": "inputs larger than L2, no flush: a step streams 151 MB in twice and 67 MB out against a 126 MB L2; the "
                             "kernels are L2-aware on purpose (the backward pass walks the rows in reverse so that it starts on "
                             "the forward pass's tail); ncu's cold-cache durations agree with the loop (profiles/)"},
            "loss": loss_val,
            # `roofline`: the STEP (forward kernel + backward kernel; 22 algorithmic bytes per row) is the
            # headline fraction, as VERDICT r1 asked; `fwd` (the dominant kernel: ~58 % of the step in the ncu launch list
            # under profiles/) and `bwd` sit beside it, each = its algorithmic bytes / its own CUDA-event time.
            # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch, read from profiles/ncu_traffic.json (the
            # summary of the committed `ncu --set full` capture of this command); null if that capture was taken at
            # another row count.
            "roofline": {"bound": "hbm", "scope": "step = forward kernel + backward kernel, 22 B/row",
                         "achieved": achieved_step, "peak": peak, "unit": "GB/s", "frac": achieved_step / peak,
                         "frac_of_8TBs": achieved_step / 8000.0, "peak_source": peak_src, "ms": ms_step,
                         "traffic": (fwd_traffic + bwd_traffic) if (fwd_traffic and bwd_traffic) else None,
                         "traffic_source": traffic_src,
                         "fwd": {"kernel": ("cox_binned_fwd_fused (9 B/row read: pass 1 + reduce + O(nbins) tail in one cooperative launch)"
                                            if fused else "cox_binned_pass1 + reduce + all-reduce + finish (9 B/row read)"),
                                 "dominant": True, "achieved": achieved_fwd, "frac": achieved_fwd / peak, "ms": fwd_ms,
                                 "traffic": fwd_traffic},
                         "bwd": {"kernel": "cox_binned_bwd (13 B/row: 9 read + 4 written)", "achieved": achieved_bwd,
                                 "frac": achieved_bwd / peak, "ms": bwd_ms, "traffic": bwd_traffic,
                                 "traffic_note": "gradient stores still in L2 when the kernel ends are not in dram__bytes_write"}},
            "e2e": {"value": e2e_val, "unit": "patients/s", "h2d_bytes_per_step": 9 * n, "d2h_bytes_per_step": 4,
                    "ms_per_step": float(e2e_t.item()) * 1e3, "api": "neg_partial_log_likelihood(log_hz, event, time) + backward, mode=auto"},
            # kernels launched by libb200surv inside the timed region on this rank, counted by the library itself
            # (b200surv_debug_launch_count): per step cox_binned_fwd_fused (cooperative) + cox_binned_bwd; with
            # --exchange nccl the forward is pass1 + reduce, the all-reduce, finish
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "parity_check": parity, "strong_scaling": strong,
            "extra": {"cindex_1m": {"n": cn, "ms": ci_ms, "patients_per_s": cn / (ci_ms * 1e-3),
                                    "ordered_pairs_per_s": sum(counts) / (ci_ms * 1e-3), "counts": counts,
                                    "n_gpus": world, "scaling": "strong (row tiles sharded, int64 all-reduce)",
                                    "algo": "1: tiled pair counting (north_star design)",
                                    "cpu_baseline": ci_cpu},
                      "cindex_1m_ranks": {"n": cn, "ms": ci2_ms, "patients_per_s": cn / (ci2_ms * 1e-3), "counts_equal_algo1": True,
                                          "n_gpus": world, "speedup_vs_pair_counting": ci_ms / ci2_ms,
                                          "algo": "2 (library default): column tiles sorted by estimate, two binary searches per "
                                                  "row and strictly-later tile, prefix subtraction around the diagonal; same "
                                                  "int64 counters bit for bit; row tiles sharded like algo 1"},
                      "cv_sweep": {"replicas_per_gpu": sw_rep, "rows_per_replica": sw_rows, "n_gpus": world,
                                   "cox_fwd_bwd_ms": sweep_cox_ms, "cindex_ms": sweep_ci_ms,
                                   "replica_evals_per_s": world * sw_rep / ((sweep_cox_ms + sweep_ci_ms) * 1e-3),
                                   "end_to_end": sweep_full,
                                   "note": "32 replicas x 100k patients per GPU packed back to back: segmented Cox "
                                           "fwd+bwd (one call) + one C-index per replica; independent replicas, no "
                                           "collective (weak scaling)"},
                      "sorted_few_ties": {"rows": n, "ms_fwd": ft_fwd_ms, "ms_bwd": ft_bwd_ms, "loss": ft_loss,
                                          "patients_per_s": n / ((ft_fwd_ms + ft_bwd_ms) * 1e-3),
                                          "frac_of_hbm_peak_22B_per_row": ALGO_BYTES_PER_ROW * n / ((ft_fwd_ms + ft_bwd_ms) * 1e-3) / 1e9 / peak,
                                          "note": "continuous times (no BINNED path): radix sort on time + look-back scans + gradient "
                                                  "scatter (mode sorted); also `bench.py --workload few_ties`; per rank"},
                      "cfg1_batch4_step": cfg1,
                      "ct_encoder": ct_extra,
                      "head_b4096": {"rows": hb, "ms_fwd_bwd": head_ms, "rows_per_s": hb / (head_ms * 1e-3),
                                     "tflops": hb * head_flop_per_row / (head_ms * 1e-3) / 1e12,
                                     "dtype": "bf16 operands, fp32 accumulate (tcgen05)", "dropout_p": 0.3,
                                     "ms_fwd_bwd_eager": head_eager_ms,
                                     "note": "gated head fwd + loss + bwd replayed as one CUDA graph (head.GraphedHeadStep), "
                                             "batch copied into the static buffers inside the timed region; "
                                             "ms_fwd_bwd_eager = the same step through the nn.Module call by call "
                                             "(host-bound: ~100 launches); per rank"}},
        }
        if cpu_val is not None:
            out["cpu_baseline"] = {"value": cpu_val, "unit": "patients/s", "cores": cores, "kind": "port",
                                   "sample": f"all {cpu_rows} rows of the same cohort, one fwd+bwd ({cpu_dt:.2f} s), torch CPU fp32 port (oracle/cox_torch.py)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default: the BASELINE 16,777,216)")
    ap.add_argument("--skip-extras", action="store_true", help="only the timed fwd+bwd loop (for ncu launch lists)")
    ap.add_argument("--workload", default="heavy_ties", choices=["heavy_ties", "few_ties"],
                    help="heavy_ties: the BASELINE configuration (integer days, BINNED path); few_ties: continuous times, SORTED path")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: per-bin sums meet inside the kernel over NVLink peer memory, or through an NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "few_ties":
        run_few_ties(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
