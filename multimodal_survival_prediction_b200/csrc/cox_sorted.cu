// Cox negative partial log-likelihood, SORTED mode: any non-negative float times, one cohort or cohorts packed back to back.
//
// The formulation BASELINE.json's north_star describes, and what the reference's own fallback does with torch ops
// (scripts/training/partial_modality_training.py:303-309: argsort + logcumsumexp): a radix sort on survival time, then
// single-pass decoupled-look-back scans over risk sets with Breslow / Efron tie handling, and a gradient scatter.  It is
// the general path behind BINNED (which needs integer day counts) and the fp64 path for hazards spread over tens of nats.
// Same math as oracle/cox.py in "sorted position" space: rows ascending in (cohort, time, events first); with tie groups
// [gs, ge) inside a cohort:
//   D = sum_{q >= gs, same cohort} w_q;  E, m = the group's event weight / event count;  l = p - gs for event rows;
//   a_p = 1 / (D - (l/m) E), f_p = (l/m) a_p;  P = sum of a over the cohort's rows up to ge - 1;  F = the group's sum of f;
//   grad = scale * (d - w (P - d F)).
// Launch sequence (all hand-written, csrc/sortscan.cuh):
//   keys    (time bits, censored bit) and the row index; per-cohort max log_hz (the exponent shift), flags
//   sort    stable LSD radix sort, 4 passes of 8 bits on the key (+ 1-2 passes on the cohort id for packed cohorts), each
//           pass = per-tile histograms, one scan of the (digit, tile) matrix, a scatter staged through shared memory
//   weights w = exp(log_hz - shift) gathered through the permutation (element-wise kernel: the gathers need occupancy)
//   R1      reverse scan: D (restarts per cohort), the
//           group-suffix sums of event weight / event count (restart per tie group: at a group's first row they are E and
//           m), ge = end of the row's group
//   F1      forward max-scan: gs = start of the row's group
//   F2      forward scan: forms a_p, f_p on the fly from (D, E, m) at gs; P (restarts per cohort), F (per group); the
//           per-cohort sums of log-denominators / event-time counts ride along (Store::finish)
//   loss    one small kernel per call: loss, scale, header of every cohort
//   grad    gradient of every row, scattered back through the permutation
// Every group sum is a SEGMENTED scan of the group's own terms -- never a difference of two running totals: with hazards
// spread over tens of nats a late group's weights are 1e-15 of the total and a difference would be rounding noise.
#include <climits>

#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

using sortscan::Tup4;

struct SegAcc {  // per cohort, device accumulators
    double sum_eta, sum_log;
    unsigned long long n_ev, n_times;
    float max_eta, max_time;
    unsigned flags, pad;
    double scale;  // d loss / d pll, written by k_loss
};

__device__ __forceinline__ uint32_t time_key(float t, bool ev) {
    return (__float_as_uint(t + 0.f) << 1) | (ev ? 0u : 1u);
}
// cohort of row / sorted position q: the cohorts are contiguous and keep their sizes under the sort
__device__ __forceinline__ int seg_of(const int64_t *__restrict__ seg_off, int n_seg, int64_t q) {
    if (seg_off == nullptr) return 0;
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg_off[mid] <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_init_acc(SegAcc *acc, int n_seg) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        SegAcc a;
        a.sum_eta = 0.0; a.sum_log = 0.0; a.n_ev = 0; a.n_times = 0;
        a.max_eta = -INFINITY; a.max_time = -INFINITY; a.flags = 0; a.pad = 0; a.scale = 0.0;
        acc[s] = a;
    }
}

// keys, row indices, cohort ids (packed cohorts only), per-cohort max log_hz / max time / flags.  One cohort: per-thread
// running values and one atomic per block; packed cohorts: one atomic per warp and iteration when the warp's 32 consecutive
// rows share a cohort (else per lane).
__global__ void __launch_bounds__(256)
k_make_keys(const float *__restrict__ log_hz, const float *__restrict__ time, const uint8_t *__restrict__ event,
            const int64_t *__restrict__ seg_off, int n_seg, int64_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
            uint32_t *__restrict__ segid, SegAcc *acc) {
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    float mx = -INFINITY, mt = -INFINITY;
    unsigned flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;  // whole warps iterate together (the shuffles below need every lane)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < n;
        const float t = in ? time[i] : 0.f, e = in ? log_hz[i] : -INFINITY;
        const unsigned bad = (in && !(t >= 0.f)) ? B200SURV_COXF_BAD_TIME : 0u;
        if (in) { keys[i] = time_key(t, event[i] != 0); vals[i] = (uint32_t)i; }
        if (seg_off == nullptr) {
            mx = fmaxf(mx, e); mt = fmaxf(mt, in ? t : -INFINITY); flags |= bad;
        } else {
            const int s = in ? seg_of(seg_off, n_seg, i) : -1;
            if (in) segid[i] = (uint32_t)s;
            const int s0 = __shfl_sync(FULL, s, 0);
            if (__all_sync(FULL, s == s0 || s < 0)) {
                const float wm = warp_max(e), wt = warp_max(in ? t : -INFINITY);
                const unsigned wf = warp_or(bad);
                if ((threadIdx.x & 31) == 0 && s0 >= 0) {
                    atomic_max_float(&acc[s0].max_eta, wm); atomic_max_float(&acc[s0].max_time, wt);
                    if (wf) atomicOr(&acc[s0].flags, wf);
                }
            } else if (in) {
                atomic_max_float(&acc[s].max_eta, e); atomic_max_float(&acc[s].max_time, t);
                if (bad) atomicOr(&acc[s].flags, bad);
            }
        }
    }
    if (seg_off == nullptr) {
        mx = block_reduce<float>(mx, -INFINITY, OpMaxF(), red_f);
        mt = block_reduce<float>(mt, -INFINITY, OpMaxF(), red_f);
        flags = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
        if (threadIdx.x == 0) {
            atomic_max_float(&acc->max_eta, mx);
            atomic_max_float(&acc->max_time, mt);
            if (flags) atomicOr(&acc->flags, flags);
        }
    }
}

// per-thread partial sums keyed by cohort, flushed with warp aggregation (Store::finish of the scans)
struct SegSums {
    int seg;
    double s0;
    long long c0;
};
template <int WHICH>  // 0: (sum_eta, n_ev)   1: (sum_log, n_times)
__device__ __forceinline__ void flush_sums(SegAcc *acc, SegSums &v) {
    if (v.seg < 0) return;
    if (WHICH == 0) { atomicAdd(&acc[v.seg].sum_eta, v.s0); atomicAdd(&acc[v.seg].n_ev, (unsigned long long)v.c0); }
    else { atomicAdd(&acc[v.seg].sum_log, v.s0); atomicAdd(&acc[v.seg].n_times, (unsigned long long)v.c0); }
    v.s0 = 0.0; v.c0 = 0;
}
template <int WHICH>
__device__ __forceinline__ void add_sums(SegAcc *acc, SegSums &v, int seg, double s, long long c) {
    if (seg != v.seg) { flush_sums<WHICH>(acc, v); v.seg = seg; }
    v.s0 += s; v.c0 += c;
}
// end of the kernel: all threads of the block arrive.  A block whose threads hold one cohort (or nothing) -- every block of
// a single-cohort call, nearly every block otherwise -- adds ONCE (two atomics per block: 16 k per scan at 16.7M rows, where
// per-warp atomics on one address would cost more than the scan); else per warp, else per lane.
template <int WHICH>
__device__ __forceinline__ void finish_sums(SegAcc *acc, SegSums &v) {
    __shared__ int sh_seg;
    __shared__ double sh_s[32];
    __shared__ long long sh_c[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (threadIdx.x == 0) sh_seg = -1;
    __syncthreads();
    if (v.seg >= 0) atomicMax(&sh_seg, v.seg);
    __syncthreads();
    const int sb = sh_seg;
    if (__syncthreads_and(v.seg == sb || v.seg < 0)) {
        const double s = warp_sum(v.seg < 0 ? 0.0 : v.s0);
        const long long c = warp_sum(v.seg < 0 ? 0ll : v.c0);
        if (lane == 0) { sh_s[warp] = s; sh_c[warp] = c; }
        __syncthreads();
        if (threadIdx.x == 0 && sb >= 0) {
            double ts = 0.0;
            long long tc = 0;
            for (int k = 0; k < nw; ++k) { ts += sh_s[k]; tc += sh_c[k]; }
            v.seg = sb; v.s0 = ts; v.c0 = tc;
            flush_sums<WHICH>(acc, v);
        }
        return;
    }
    int s0 = v.seg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s0 = max(s0, __shfl_xor_sync(FULL, s0, o));  // the cohort of the lanes that hold one
    if (__all_sync(FULL, v.seg == s0 || v.seg < 0)) {
        const double s = warp_sum(v.seg < 0 ? 0.0 : v.s0);
        const long long c = warp_sum(v.seg < 0 ? 0ll : v.c0);
        if (lane == 0 && s0 >= 0) { v.seg = s0; v.s0 = s; v.c0 = c; flush_sums<WHICH>(acc, v); }
    } else {
        flush_sums<WHICH>(acc, v);
    }
}

// ---- weights in sorted order (a plain element-wise kernel: the two dependent gathers and the fp64 exp need the occupancy
// a 148-register scan kernel does not have -- fused into the scan's load they cost 240 us per 4M rows at 12 % warp
// occupancy, ncu r2_segscan); per-cohort sum of the event rows' log_hz and event count ride along
__global__ void __launch_bounds__(256)
k_weights(const float *__restrict__ log_hz, const uint32_t *__restrict__ keys_s, const uint32_t *__restrict__ idx_s,
          const int64_t *__restrict__ seg_off, int n_seg, int64_t n, SegAcc *acc, float *__restrict__ w) {
    __shared__ double red_d[32];
    __shared__ long long red_l[32];
    double se = 0.0;
    long long ne = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, n_round = (n + 31) / 32 * 32;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
        const bool in = p < n;
        const int s = in ? seg_of(seg_off, n_seg, p) : -1;
        const float eta = in ? log_hz[idx_s[p]] : 0.f;
        const bool d = in && !(keys_s[p] & 1u);
        if (in) w[p] = (float)exp((double)eta - (double)acc[s].max_eta);
        if (seg_off == nullptr) {
            if (d) { se += (double)eta; ne += 1; }
        } else {  // packed cohorts: one pair of atomics per warp and iteration when its 32 positions share a cohort
            const int s0 = __shfl_sync(FULL, s, 0);
            if (__all_sync(FULL, s == s0 || s < 0)) {
                const double ws = warp_sum(d ? (double)eta : 0.0);
                const long long wc = warp_sum(d ? 1ll : 0ll);
                if ((threadIdx.x & 31) == 0 && s0 >= 0 && wc > 0) {
                    atomicAdd(&acc[s0].sum_eta, ws); atomicAdd(&acc[s0].n_ev, (unsigned long long)wc);
                }
            } else if (d) {
                atomicAdd(&acc[s].sum_eta, (double)eta); atomicAdd(&acc[s].n_ev, 1ull);
            }
        }
    }
    if (seg_off == nullptr) {
        se = block_reduce<double>(se, 0.0, OpAddD(), red_d);
        ne = block_reduce<long long>(ne, 0ll, OpAddLL(), red_l);
        if (threadIdx.x == 0 && ne > 0) { atomicAdd(&acc->sum_eta, se); atomicAdd(&acc->n_ev, (unsigned long long)ne); }
    }
}

// ---- R1 (reverse): D, group-suffix sums, group ends
struct LoadR1 {
    const float *w; const uint32_t *keys_s; const int64_t *seg_off; int n_seg; int64_t n;
    __device__ Tup4 operator()(int64_t p) const {
        Tup4 t;
        const uint32_t k = keys_s[p];
        const int s = seg_of(seg_off, n_seg, p);
        const bool seg_tail = seg_off ? (p + 1 == seg_off[s + 1]) : (p == n - 1);
        const bool tail = seg_tail || ((keys_s[p + 1] >> 1) != (k >> 1));
        const double w = (double)this->w[p];
        const double d = (k & 1u) ? 0.0 : 1.0;
        t.a = w; t.b = w * d; t.c = d;
        t.i = (tail ? ((p + 2) | sortscan::T4_GROUP) : 0) | (seg_tail ? sortscan::T4_SEG : 0);
        return t;
    }
};
struct StoreR1 {
    double *D, *Esuf; int *msuf, *ge;
    __device__ void operator()(int64_t p, const Tup4 &inc, const Tup4 &) {
        D[p] = inc.a; Esuf[p] = inc.b; msuf[p] = (int)(inc.c + 0.5);
        ge[p] = (int)((inc.i & sortscan::T4_POS) - 1);
    }
    __device__ void finish() {}
};
// ---- F1 (forward): group starts
struct LoadF1 {
    const uint32_t *keys_s; const int64_t *seg_off; int n_seg;
    __device__ Tup4 operator()(int64_t p) const {
        Tup4 t = sortscan::t4_identity();
        const int s = seg_of(seg_off, n_seg, p);
        const bool seg_head = seg_off ? (p == seg_off[s]) : (p == 0);
        const bool head = seg_head || ((keys_s[p - 1] >> 1) != (keys_s[p] >> 1));
        t.i = (head ? ((p + 1) | sortscan::T4_GROUP) : 0) | (seg_head ? sortscan::T4_SEG : 0);
        return t;
    }
};
struct StoreF1 {
    int *gs;
    __device__ void operator()(int64_t p, const Tup4 &inc, const Tup4 &) { gs[p] = (int)((inc.i & sortscan::T4_POS) - 1); }
    __device__ void finish() {}
};
// ---- F2 (forward): a_p, f_p on the fly; P, F; per-cohort log-denominator sums and event-time counts
struct LoadF2 {
    const uint32_t *keys_s; const int *gs; const double *D, *Esuf; const int *msuf; const int64_t *seg_off; int n_seg; int efron;
    __device__ Tup4 operator()(int64_t p) const {
        Tup4 t = sortscan::t4_identity();
        const int g0 = gs[p];
        const int s = seg_of(seg_off, n_seg, p);
        const bool seg_head = seg_off ? (p == seg_off[s]) : (p == 0);
        t.i = (g0 == (int)p ? sortscan::T4_GROUP : 0) | (seg_head ? sortscan::T4_SEG : 0);
        if (!(keys_s[p] & 1u)) {
            double den = D[g0], frac = 0.0;
            if (efron) { frac = (double)((int)p - g0) / (double)msuf[g0]; den -= frac * Esuf[g0]; }
            t.a = 1.0 / den; t.b = frac / den;
        }
        return t;
    }
};
struct StoreF2 {
    const uint32_t *keys_s; const int *gs; const int64_t *seg_off; int n_seg; SegAcc *acc;
    double *PA, *PF;
    SegSums sums;
    __device__ void operator()(int64_t p, const Tup4 &inc, const Tup4 &el) {
        PA[p] = inc.a; PF[p] = inc.b;
        if (!(keys_s[p] & 1u)) {  // event row: log(den) = -log(a_p); one event time per group (counted at its first row)
            const int s = seg_of(seg_off, n_seg, p);
            add_sums<1>(acc, sums, s, -log(el.a) + (double)acc[s].max_eta, gs[p] == (int)p ? 1 : 0);
        }
    }
    __device__ void finish() { finish_sums<1>(acc, sums); }
};

__global__ void __launch_bounds__(256)
k_loss(SegAcc *acc, int n_seg, int ties, int reduction, float *__restrict__ out_loss, b200surv_cox_header *__restrict__ hdrs) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        SegAcc &a = acc[s];
        const double pll = a.sum_eta - a.sum_log;
        const double n_ev = (double)a.n_ev, n_times = (double)a.n_times;
        double norm = 1.0;
        if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = n_ev;
        else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = (ties == B200SURV_TIES_EFRON) ? n_times : n_ev;
        double scale = a.n_ev > 0 ? -1.0 / norm : 0.0, loss = a.n_ev > 0 ? -pll / norm : 0.0;
        if (a.flags) { loss = __longlong_as_double(0x7ff8000000000000ll); scale = loss; }
        a.scale = scale;
        b200surv_cox_header *hdr = hdrs + s;
        hdr->flags = a.flags; hdr->mode = B200SURV_COX_SORTED; hdr->loss = (float)loss; hdr->scale = (float)scale;
        hdr->shift = a.max_eta; hdr->max_log_hz = a.max_eta; hdr->max_time = a.max_time; hdr->nbins = 0;
        hdr->n_events = (int64_t)a.n_ev; hdr->n_event_times = (int64_t)a.n_times; hdr->pll = pll; hdr->min_log_hz = 0.f;
        hdr->reserved = 0;
        out_loss[s] = (float)loss;
    }
}

// the state keeps the UNSCALED per-row gradient d loss / d log_hz for grad_out = 1 (cox_scale_grad multiplies by grad_out)
__global__ void __launch_bounds__(256)
k_grad(const uint32_t *__restrict__ keys_s, const uint32_t *__restrict__ idx_s, int64_t n, const float *__restrict__ w,
       const double *__restrict__ PA, const double *__restrict__ PF, const int *__restrict__ ge, const int64_t *__restrict__ seg_off,
       int n_seg, const SegAcc *__restrict__ acc, float *__restrict__ grad_unit) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double d = (keys_s[p] & 1u) ? 0.0 : 1.0;
        const int gend = ge[p];
        const double g = d - (double)w[p] * (PA[gend - 1] - d * PF[gend - 1]);
        grad_unit[idx_s[p]] = (float)(acc[seg_of(seg_off, n_seg, p)].scale * g);
    }
}

struct SortedLayout {
    size_t off_acc, off_keys, off_vals, off_keys_s, off_idx_s, off_segid, off_w, off_D, off_E, off_m, off_ge, off_gs, off_PA, off_PF,
        off_tmp, total;
};

SortedLayout sorted_layout(int64_t n, int64_t n_seg) {
    SortedLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.off_acc = take((size_t)n_seg * sizeof(SegAcc));
    L.off_keys = take(N * 4 + 4); L.off_vals = take(N * 4); L.off_keys_s = take(N * 4 + 4); L.off_idx_s = take(N * 4);
    L.off_segid = take(n_seg > 1 ? N * 4 : 4);
    L.off_w = take(N * 4); L.off_D = take(N * 8); L.off_E = take(N * 8); L.off_m = take(N * 4); L.off_ge = take(N * 4);
    L.off_gs = take(N * 4); L.off_PA = take(N * 8); L.off_PF = take(N * 8);
    size_t tmp = sortscan::radix_sort_temp_bytes((int64_t)N), sc = sortscan::seg_scan_state_bytes((int64_t)N);
    L.off_tmp = take(tmp > sc ? tmp : sc);
    L.total = o;
    return L;
}

}  // namespace

size_t cox_sorted_workspace_bytes(int64_t n, int64_t n_seg) { return sorted_layout(n, n_seg < 1 ? 1 : n_seg).total; }

int32_t cox_sorted_fwd_launch(const float *log_hz, const float *time, const uint8_t *event, const int64_t *seg_off, int64_t n,
                              int64_t n_seg, int ties, int reduction, float *out_loss, void *state, size_t state_bytes,
                              void *ws, size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(n >= 1 && n < (int64_t)INT_MAX - 2, "n must be in [1, 2^31)");
    B200_REQUIRE(n_seg >= 1 && n_seg <= 65535, "n_seg must be in [1, 65535]");
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    if (n_seg == 1) seg_off = nullptr;
    const SortedLayout L = sorted_layout(n, n_seg);
    if (ws_bytes < L.total) { set_error("cox sorted: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    const size_t need = (size_t)n_seg * sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
    if (state_bytes < need) { set_error("cox sorted: state buffer %zu < %zu", state_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    SegAcc *acc = reinterpret_cast<SegAcc *>(w8 + L.off_acc);
    uint32_t *keys = reinterpret_cast<uint32_t *>(w8 + L.off_keys), *vals = reinterpret_cast<uint32_t *>(w8 + L.off_vals);
    uint32_t *keys_s = reinterpret_cast<uint32_t *>(w8 + L.off_keys_s), *idx_s = reinterpret_cast<uint32_t *>(w8 + L.off_idx_s);
    uint32_t *segid = reinterpret_cast<uint32_t *>(w8 + L.off_segid);
    float *wv = reinterpret_cast<float *>(w8 + L.off_w);
    double *Dv = reinterpret_cast<double *>(w8 + L.off_D), *Ev = reinterpret_cast<double *>(w8 + L.off_E);
    int *mv = reinterpret_cast<int *>(w8 + L.off_m), *ge = reinterpret_cast<int *>(w8 + L.off_ge), *gs = reinterpret_cast<int *>(w8 + L.off_gs);
    double *PA = reinterpret_cast<double *>(w8 + L.off_PA), *PF = reinterpret_cast<double *>(w8 + L.off_PF);
    void *tmp = w8 + L.off_tmp;
    int grid = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    if (grid > cap) grid = cap;
    const int nseg = (int)n_seg;

    b200surv_cox_header *hdrs = static_cast<b200surv_cox_header *>(state);
    float *grad_unit = reinterpret_cast<float *>(hdrs + n_seg);
    int32_t rc;

    k_init_acc<<<(nseg + 255) / 256, 256, 0, st>>>(acc, nseg);
    // keys are generated into (keys_s, idx_s); radix_sort_pairs2 reports which buffer pair holds the result
    k_make_keys<<<grid, 256, 0, st>>>(log_hz, time, event, seg_off, nseg, n, keys_s, idx_s, segid, acc);
    const int seg_bits = n_seg == 1 ? 0 : (n_seg <= 256 ? 8 : 16);
    int in_first = 1;
    rc = sortscan::radix_sort_pairs2(keys_s, idx_s, keys, vals, n, 32, n_seg > 1 ? segid : nullptr, seg_bits, tmp, st, &in_first);
    if (rc) return rc;
    const uint32_t *ks = in_first ? keys_s : keys, *is = in_first ? idx_s : vals;
    k_weights<<<grid, 256, 0, st>>>(log_hz, ks, is, seg_off, nseg, n, acc, wv);
    rc = sortscan::seg_scan<sortscan::P_MIN, true, 2, 1, 1>(n, LoadR1{wv, ks, seg_off, nseg, n}, StoreR1{Dv, Ev, mv, ge}, tmp, st);
    if (rc) return rc;
    rc = sortscan::seg_scan<sortscan::P_MAX, false, 0, 0, 0>(n, LoadF1{ks, seg_off, nseg}, StoreF1{gs}, tmp, st);
    if (rc) return rc;
    rc = sortscan::seg_scan<sortscan::P_NONE, false, 2, 1, 0>(
        n, LoadF2{ks, gs, Dv, Ev, mv, seg_off, nseg, ties == B200SURV_TIES_EFRON ? 1 : 0},
        StoreF2{ks, gs, seg_off, nseg, acc, PA, PF, {-1, 0.0, 0}}, tmp, st);
    if (rc) return rc;
    k_loss<<<(nseg + 255) / 256, 256, 0, st>>>(acc, nseg, ties, reduction, out_loss, hdrs);
    k_grad<<<grid, 256, 0, st>>>(ks, is, n, wv, PA, PF, ge, seg_off, nseg, acc, grad_unit);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(2 + 4 * (4 + seg_bits / 8) + 6 + 2);
    return B200SURV_OK;
}

}  // namespace b200surv
