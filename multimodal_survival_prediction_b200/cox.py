"""Cox negative partial log-likelihood on B200 -- host-side mirror of torchsurv's operator.

Same call signature as ``torchsurv.loss.cox.neg_partial_log_likelihood(log_hz, event, time,
ties_method="efron", reduction="mean", checks=True)`` -- the function the reference calls with three
positional arguments at scripts/training/partial_modality_training.py:285-288,
simple_fusion.py:270,311 and final_multimodal.py:158-162.  All arithmetic runs in libb200surv.so
(csrc/cox_*.cu) through the C ABI of include/b200surv.h; there is no CPU path: CPU tensors are copied
to the current CUDA device and the result is returned on the input's device.

Conventions that cannot be verified against torchsurv in this image are explicit
keyword arguments whose defaults follow the recollected torchsurv behaviour:
``efron_mean_over="event_times"`` (torchsurv averages Efron's one-term-per-distinct-event-time vector;
``"events"`` divides by the event count like the reference's in-repo fallback,
partial_modality_training.py:309).
"""
from __future__ import annotations

import ctypes
import warnings

import torch

from . import _lib as L

_MODES = {"auto": 0, "small": L.COX_SMALL, "binned": L.COX_BINNED, "sorted": L.COX_SORTED}
DEFAULT_NBINS = 4096


def _reduction_code(ties_method: str, reduction: str, efron_mean_over: str) -> int:
    r = reduction.lower()
    if r == "sum":
        return L.REDUCE_SUM
    if r != "mean":
        raise ValueError(f'Reduction {reduction} is not implemented yet, should be one of ["mean", "sum"].')
    if efron_mean_over == "events":
        return L.REDUCE_MEAN_EVENTS
    if efron_mean_over != "event_times":
        raise ValueError("efron_mean_over must be 'event_times' or 'events'")
    return L.REDUCE_MEAN_TERMS


def _ties_code(ties_method: str) -> int:
    if ties_method not in L.TIES:
        raise ValueError(f'Ties method {ties_method} should be one of ["efron", "breslow"]')
    return L.TIES[ties_method]


def _validate(log_hz, event, time):
    for name, t in (("log_hz", log_hz), ("event", event), ("time", time)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"Input '{name}' should be a tensor")
    if event.dtype != torch.bool:
        raise ValueError("Input 'event' should be of boolean type (use event.bool())")
    if not torch.is_floating_point(log_hz):
        raise ValueError("Input 'log_hz' should be of float type")
    if not torch.is_floating_point(time):
        raise ValueError("Input 'time' should be of float type")
    if event.dim() != 1 or time.dim() != 1:
        raise ValueError("Inputs 'event' and 'time' should be one-dimensional")
    if log_hz.dim() == 2 and log_hz.shape[1] == 1:
        pass
    elif log_hz.dim() != 1:
        raise ValueError("Input 'log_hz' should have shape (n,) or (n, 1)")
    if not (log_hz.shape[0] == event.shape[0] == time.shape[0]):
        raise ValueError("Dimension mismatch: 'log_hz', 'event' and 'time' must have the same length")


def _device_for(*tensors):
    for t in tensors:
        if t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise L.B200SurvError("no CUDA device: the B200 survival kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def read_headers(state: torch.Tensor, n_seg: int, mode: int = L.COX_SMALL):
    """Device->host copy of the per-segment headers (synchronises the stream).

    SMALL / SORTED keep the n_seg headers contiguous at the start of ``state``; BINNED interleaves them with the
    (P, F) tables: segment s's header starts at ``s * state.numel() // n_seg`` (include/b200surv.h)."""
    if mode == L.COX_BINNED and n_seg > 1:
        stride = state.numel() // n_seg
        raw = state.view(n_seg, stride)[:, : L.COX_HEADER_BYTES].contiguous().cpu().numpy().tobytes()
    else:
        raw = state[: n_seg * L.COX_HEADER_BYTES].cpu().numpy().tobytes()
    return [L.CoxHeader.from_buffer_copy(raw, i * L.COX_HEADER_BYTES) for i in range(n_seg)]


def cox_fwd_raw(log_hz, time, event, seg_offsets, n_seg, ties, reduction, mode, nbins, shift=0.0, state=None,
                loss=None):
    """One call of b200surv_cox_fwd on prepared device tensors.  Returns (loss[n_seg], state); ``state`` / ``loss``
    may be caller-provided views (re-running one cohort of a packed set in place)."""
    dev = log_hz.device
    L.require_device(dev.index)
    lib = L.load()
    n = log_hz.numel()
    sb = lib.b200surv_cox_state_bytes(n, n_seg, mode, nbins)
    wb = lib.b200surv_cox_workspace_bytes(n, n_seg, mode, nbins)
    if state is None:
        state = torch.empty(sb, dtype=torch.uint8, device=dev)
    ws = torch.empty(max(wb, 256), dtype=torch.uint8, device=dev)
    if loss is None:
        loss = torch.empty(n_seg, dtype=torch.float32, device=dev)
    rc = lib.b200surv_cox_fwd(L.ptr(log_hz), L.ptr(time), L.ptr(event), L.ptr(seg_offsets), n, n_seg, ties,
                              reduction, mode, nbins, ctypes.c_float(shift), L.ptr(loss), L.ptr(state), sb,
                              L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "b200surv_cox_fwd")
    return loss, state


def cox_bwd_raw(grad_out, state, log_hz, time, event, seg_offsets, n_seg, mode, nbins):
    dev = log_hz.device
    lib = L.load()
    n = log_hz.numel()
    out = torch.empty(n, dtype=torch.float32, device=dev)
    rc = lib.b200surv_cox_bwd(L.ptr(grad_out), L.ptr(state), state.numel(), L.ptr(log_hz), L.ptr(time),
                              L.ptr(event), L.ptr(seg_offsets), n, n_seg, mode, nbins, L.ptr(out),
                              L.stream_ptr(dev))
    L.check(rc, "b200surv_cox_bwd")
    return out


LOWP_MIN = -11.090354888959125      # -16 ln 2: B200SURV_COXF_LOW_PRECISION threshold on min(log_hz) - shift


def _fit_shift(n, hi, lo):
    """Exponent shift for the 36.28 fixed point of a cohort of n rows with log_hz in [lo, hi]: the largest weight is 2^k
    with n * 2^k <= 2^29 (full precision without overflow).  None when the smallest weight would still fall below the
    LOW_PRECISION floor: the spread does not fit at any shift."""
    k = min(28, max(0, 29 - max(1, (n - 1).bit_length())))
    shift = hi - k * 0.6931471805599453
    return shift if lo - shift >= LOWP_MIN else None


def _plan_and_run(log_hz, time, event, seg_offsets, n_seg, max_seg, ties, reduction, mode, nbins, checks=True):
    """Mode policy.  Explicit modes never synchronise (unless ``checks`` asks for the SMALL header); "auto" reads the
    64-byte headers back once (torchsurv itself synchronises on event.sum() and torch.unique)."""
    n = log_hz.numel()
    if mode != 0:
        nb = nbins or DEFAULT_NBINS
        loss, state = cox_fwd_raw(log_hz, time, event, seg_offsets, n_seg, ties, reduction, mode, nb)
        return loss, state, mode, nb
    if max_seg <= L.COX_SMALL_MAX:
        loss, state = cox_fwd_raw(log_hz, time, event, seg_offsets, n_seg, ties, reduction, L.COX_SMALL, 0)
        if checks:      # like the larger cohorts: bad times raise instead of returning a silent NaN
            if any(h.flags & L.COXF_BAD_TIME for h in read_headers(state, n_seg, L.COX_SMALL)):
                raise ValueError("Input 'time' should be non-negative and free of NaN")
        return loss, state, L.COX_SMALL, 0
    nb = nbins or DEFAULT_NBINS
    shift = 0.0
    for _attempt in range(4):
        loss, state = cox_fwd_raw(log_hz, time, event, seg_offsets, n_seg, ties, reduction, L.COX_BINNED, nb,
                                  shift)
        hdrs = read_headers(state, n_seg, L.COX_BINNED)
        flags = 0
        for h in hdrs:
            flags |= h.flags
        if flags == 0:
            return loss, state, L.COX_BINNED, nb
        if flags & L.COXF_BAD_TIME:
            raise ValueError("Input 'time' should be non-negative and free of NaN")
        if flags & L.COXF_NOT_BINNABLE:
            if nbins is None and nb < L.COX_MAX_BINS:   # maybe integer days beyond 4095: one wider try
                nb = L.COX_MAX_BINS
                continue
            break
        if flags & (L.COXF_EXP_RANGE | L.COXF_LOW_PRECISION):
            if n_seg > 1:
                # packed cohorts: each flagged cohort is re-run in place with ITS OWN shift (the state keeps one header
                # and table per cohort, and the backward pass reads the shift from the cohort's header)
                so_host = seg_offsets.cpu().tolist()
                stride = state.numel() // n_seg
                refit = True
                for s_, h in enumerate(hdrs):
                    if not h.flags:
                        continue
                    a, b = so_host[s_], so_host[s_ + 1]
                    sh = _fit_shift(b - a, h.max_log_hz, h.min_log_hz)
                    if sh is None:      # this cohort's spread fits no shift: the whole set goes to the fp64 path
                        refit = False
                        break
                    cox_fwd_raw(log_hz[a:b], time[a:b], event[a:b], None, 1, ties, reduction, L.COX_BINNED, nb, sh,
                                state=state[s_ * stride:(s_ + 1) * stride], loss=loss[s_:s_ + 1])
                if not refit or any(h.flags for h in read_headers(state, n_seg, L.COX_BINNED)):
                    break
                return loss, state, L.COX_BINNED, nb
            new_shift = _fit_shift(n, max(h.max_log_hz for h in hdrs), min(h.min_log_hz for h in hdrs))
            if new_shift is None or new_shift == shift:
                break       # the spread of log_hz does not fit the fixed point at any shift: fp64 path
            shift = new_shift
            continue
    # any non-negative float times, hazards of any spread, one cohort or packed cohorts: sort + scans in fp64
    loss, state = cox_fwd_raw(log_hz, time, event, seg_offsets if n_seg > 1 else None, n_seg, ties, reduction, L.COX_SORTED, 0)
    if checks and any(h.flags & L.COXF_BAD_TIME for h in read_headers(state, n_seg, L.COX_SORTED)):
        raise ValueError("Input 'time' should be non-negative and free of NaN")
    return loss, state, L.COX_SORTED, 0


class _CoxNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_hz, event, time, seg_offsets, n_seg, max_seg, ties, reduction, mode, nbins, checks=True):
        dev = _device_for(log_hz, time, event)
        x = log_hz.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        t = time.detach().to(device=dev, dtype=torch.float32).contiguous()
        e = event.detach().to(device=dev).contiguous()
        so = None if seg_offsets is None else seg_offsets.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            loss, state, used_mode, nb = _plan_and_run(x, t, e, so, n_seg, max_seg, ties, reduction, mode, nbins,
                                                       checks)
        ctx.save_for_backward(x, t, e, state, so if so is not None else torch.empty(0, device=dev))
        ctx.meta = (n_seg, used_mode, nb, log_hz.shape, log_hz.dtype, log_hz.device, so is not None)
        return loss.to(log_hz.device)

    @staticmethod
    def backward(ctx, grad_out):
        x, t, e, state, so = ctx.saved_tensors
        n_seg, mode, nb, shape, dtype, src_dev, has_so = ctx.meta
        g = grad_out.detach().reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(x.device):
            grad = cox_bwd_raw(g, state, x, t, e, so if has_so else None, n_seg, mode, nb)
        grad = grad.reshape(shape).to(device=src_dev, dtype=dtype)
        return grad, None, None, None, None, None, None, None, None, None, None


def neg_partial_log_likelihood(log_hz, event, time, ties_method="efron", reduction="mean", checks=True, *,
                               efron_mean_over="event_times", mode="auto", nbins=None):
    """Negative Cox partial log-likelihood (0-dim fp32 tensor carrying grad to ``log_hz``).

    log_hz: (n,) or (n,1) float; event: (n,) bool; time: (n,) float, non-negative.
    mode: "auto" | "small" | "binned" | "sorted" (see include/b200surv.h)."""
    if checks:
        _validate(log_hz, event, time)
    n = event.shape[0]
    if n == 0:
        warnings.warn("No events OR single sample. Returning zero loss for the batch")
        return log_hz.sum() * 0.0
    ties = _ties_code(ties_method)
    red = _reduction_code(ties_method, reduction, efron_mean_over)
    loss = _CoxNLL.apply(log_hz, event, time, None, 1, n, ties, red, _MODES[mode], nbins, bool(checks))
    return loss.reshape(())


def neg_partial_log_likelihood_segmented(log_hz, event, time, seg_offsets, ties_method="efron",
                                         reduction="mean", checks=True, *, efron_mean_over="event_times",
                                         mode="auto", nbins=None):
    """Independent cohorts packed back to back (CV sweep, BASELINE.json configs[4]).

    seg_offsets: int64 tensor [n_seg+1] (host or device), seg_offsets[0] == 0, last == n.
    Returns a (n_seg,) tensor of losses."""
    if checks:
        _validate(log_hz, event, time)
    so_host = seg_offsets.detach().cpu().to(torch.int64)
    n_seg = so_host.numel() - 1
    if n_seg < 1 or int(so_host[0]) != 0 or int(so_host[-1]) != event.shape[0]:
        raise ValueError("seg_offsets must start at 0 and end at n")
    lens = so_host[1:] - so_host[:-1]
    if bool((lens <= 0).any()):
        raise ValueError("empty segments are not allowed")
    ties = _ties_code(ties_method)
    red = _reduction_code(ties_method, reduction, efron_mean_over)
    return _CoxNLL.apply(log_hz, event, time, so_host, n_seg, int(lens.max()), ties, red, _MODES[mode], nbins,
                         bool(checks))
