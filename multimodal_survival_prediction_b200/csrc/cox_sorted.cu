// Cox negative partial log-likelihood, SORTED mode: any non-negative float times, one cohort.
//
// General fallback behind BINNED (which needs integer day counts).  Same math as oracle/cox.py in
// "sorted position" space: rows ascending in (time, events first); with tie groups [gs, ge):
//   D = sum_{q >= gs} w_q,  E, m = segmented sums over the group,  l = p - gs for event rows,
//   a_p = 1/(D - (l/m)E), f_p = (l/m) a_p,  P = prefix sum of a up to ge-1,  F = segmented sum of f,
//   grad = scale * (d - w (P - d F)).
// The radix sort and the three device-wide scans are hand-written (sortscan.cuh): a stable LSD radix sort on
// (time, event) keys and single-pass scans with decoupled look-back --
//   R1 (reverse): D = suffix sums of w, ge = end of the row's tie group (min-scan)
//   F1 (forward): SEGMENTED sums (restarting at every tie group) of the event weights and of the event count, gs =
//                 start of the tie group (max-scan); a group's E and m are the values at its last row
//   F2 (forward): prefix sums of a; segmented sums of f (F of a group = the value at its last row)
// Group sums are segmented scans, not differences of global prefix sums: with log-hazards spread over tens of nats
// (the cohorts BINNED hands over, COXF_LOW_PRECISION) a late group's weights are 1e-15 of the running total and a
// difference would be pure rounding noise (negative Efron denominators, NaN).
#include <climits>

#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

struct Acc {  // device accumulators
    double sum_eta, sum_log;
    unsigned long long n_ev, n_times;
    float max_eta, max_time;
    unsigned flags, pad;
};

__device__ __forceinline__ uint32_t time_key(float t, bool ev) {
    return (__float_as_uint(t + 0.f) << 1) | (ev ? 0u : 1u);
}

__global__ void __launch_bounds__(256)
k_init_acc(Acc *acc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        acc->sum_eta = 0.0; acc->sum_log = 0.0; acc->n_ev = 0; acc->n_times = 0;
        acc->max_eta = -INFINITY; acc->max_time = -INFINITY; acc->flags = 0; acc->pad = 0;
    }
}

__global__ void __launch_bounds__(256)
k_make_keys(const float *__restrict__ log_hz, const float *__restrict__ time,
            const uint8_t *__restrict__ event, int64_t n, uint32_t *__restrict__ keys,
            uint32_t *__restrict__ vals, Acc *acc) {
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    float mx = -INFINITY, mt = -INFINITY;
    unsigned flags = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float t = time[i];
        if (!(t >= 0.f)) flags |= B200SURV_COXF_BAD_TIME;
        keys[i] = time_key(t, event[i] != 0);
        vals[i] = (uint32_t)i;
        mx = fmaxf(mx, log_hz[i]);
        mt = fmaxf(mt, t);
    }
    mx = block_reduce<float>(mx, -INFINITY, OpMaxF(), red_f);
    mt = block_reduce<float>(mt, -INFINITY, OpMaxF(), red_f);
    flags = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
    if (threadIdx.x == 0) {
        atomic_max_float(&acc->max_eta, mx);
        atomic_max_float(&acc->max_time, mt);
        if (flags) atomicOr(&acc->flags, flags);
    }
}

// sorted rows: weights
__global__ void __launch_bounds__(256)
k_gather(const float *__restrict__ log_hz, const uint32_t *__restrict__ idx_s, int64_t n, const Acc *__restrict__ acc,
         double *__restrict__ w) {
    const double c = (double)acc->max_eta;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
        w[p] = exp((double)log_hz[idx_s[p]] - c);
}

// scan functors (see sortscan.cuh: Tup = {a, b: fp64 sums; i: integer with add / min / max})
struct LoadR1 {   // a = w, i = p + 1 at the last row of a tie group
    const double *w; const uint32_t *keys_s; int64_t n;
    __device__ sortscan::Tup operator()(int64_t p) const {
        sortscan::Tup t;
        t.a = w[p]; t.b = 0.0;
        const bool tail = (p == n - 1) || ((keys_s[p + 1] >> 1) != (keys_s[p] >> 1));
        t.i = tail ? p + 1 : LLONG_MAX;
        return t;
    }
};
struct StoreR1 {
    double *D; int *ge;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &) const { D[p] = inc.a; ge[p] = (int)inc.i; }
};
struct LoadF1 {   // a = event weight, b = event indicator, i = p at the first row of a tie group
    const double *w; const uint32_t *keys_s;
    __device__ sortscan::Tup operator()(int64_t p) const {
        sortscan::Tup t;
        const uint32_t k = keys_s[p];
        const bool d = !(k & 1u);
        t.a = d ? w[p] : 0.0; t.b = d ? 1.0 : 0.0;
        const bool head = (p == 0) || ((keys_s[p - 1] >> 1) != (k >> 1));
        t.i = head ? p : -1;
        return t;
    }
};
struct StoreF1 {
    double *prefE, *cntE; int *gs;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &) const {
        prefE[p] = inc.a; cntE[p] = inc.b; gs[p] = (int)inc.i;
    }
};
struct LoadF2 {   // a = a_p (global prefix), b = f_p (segmented by tie group: head marker i = p)
    const double *a, *f; const int *gs;
    __device__ sortscan::Tup operator()(int64_t p) const {
        sortscan::Tup t; t.a = a[p]; t.b = f[p]; t.i = gs[p] == (int)p ? p : -1; return t;
    }
};
struct StoreF2 {
    double *PA, *PF;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &) const { PA[p] = inc.a; PF[p] = inc.b; }
};

__global__ void __launch_bounds__(256)
k_terms(const float *__restrict__ log_hz, const uint32_t *__restrict__ keys_s,
        const uint32_t *__restrict__ idx_s, int64_t n, int ties, const double *__restrict__ Dpos,
        const double *__restrict__ prefE, const double *__restrict__ cntE, const int *__restrict__ gsv,
        const int *__restrict__ ge, Acc *acc, double *__restrict__ a, double *__restrict__ f) {
    __shared__ double red_d[32];
    __shared__ long long red_l[32];
    const double c = (double)acc->max_eta;
    double sum_eta = 0.0, sum_log = 0.0;
    long long n_ev = 0, n_times = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        double ap = 0.0, fp = 0.0;
        if (!(keys_s[p] & 1u)) {
            const int gs = gsv[p], gend = ge[p];
            const double D = Dpos[gs];
            const int l = (int)p - gs;
            double den = D, frac = 0.0;
            if (ties == B200SURV_TIES_EFRON) {
                const double E = prefE[gend - 1], m = cntE[gend - 1];   // segmented sums: the group's own terms only
                frac = (double)l / m;
                den = D - frac * E;
            }
            ap = 1.0 / den;
            fp = frac / den;
            sum_log += log(den) + c;
            sum_eta += (double)log_hz[idx_s[p]];
            n_ev += 1;
            n_times += (l == 0);
        }
        a[p] = ap;
        f[p] = fp;
    }
    sum_eta = block_reduce<double>(sum_eta, 0.0, OpAddD(), red_d);
    sum_log = block_reduce<double>(sum_log, 0.0, OpAddD(), red_d);
    n_ev = block_reduce<long long>(n_ev, 0ll, OpAddLL(), red_l);
    n_times = block_reduce<long long>(n_times, 0ll, OpAddLL(), red_l);
    if (threadIdx.x == 0) {
        atomicAdd(&acc->sum_eta, sum_eta);
        atomicAdd(&acc->sum_log, sum_log);
        atomicAdd(&acc->n_ev, (unsigned long long)n_ev);
        atomicAdd(&acc->n_times, (unsigned long long)n_times);
    }
}

__device__ __forceinline__ void loss_from_acc(const Acc *acc, int ties, int reduction, double *loss,
                                              double *scale, double *pll_out) {
    const double pll = acc->sum_eta - acc->sum_log;
    const double n_ev = (double)acc->n_ev, n_times = (double)acc->n_times;
    double norm = 1.0;
    if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = n_ev;
    else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = (ties == B200SURV_TIES_EFRON) ? n_times : n_ev;
    *scale = acc->n_ev > 0 ? -1.0 / norm : 0.0;
    *loss = acc->n_ev > 0 ? -pll / norm : 0.0;
    if (acc->flags) { *loss = __longlong_as_double(0x7ff8000000000000ll); *scale = *loss; }
    *pll_out = pll;
}

__global__ void __launch_bounds__(256)
k_grad(const uint32_t *__restrict__ keys_s, const uint32_t *__restrict__ idx_s, int64_t n, int ties,
       int reduction, const double *__restrict__ w, const double *__restrict__ PA,
       const double *__restrict__ PF, const int *__restrict__ gsv, const int *__restrict__ ge, const Acc *__restrict__ acc,
       float *__restrict__ grad_unit, float *__restrict__ out_loss, b200surv_cox_header *hdr) {
    double loss, scale, pll;
    loss_from_acc(acc, ties, reduction, &loss, &scale, &pll);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double d = (keys_s[p] & 1u) ? 0.0 : 1.0;
        const int gs = gsv[p], gend = ge[p];
        const double F = PF[gend - 1];
        const double g = d - w[p] * (PA[gend - 1] - d * F);
        grad_unit[idx_s[p]] = (float)(scale * g);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hdr->flags = acc->flags; hdr->mode = B200SURV_COX_SORTED; hdr->loss = (float)loss;
        hdr->scale = (float)scale; hdr->shift = acc->max_eta; hdr->max_log_hz = acc->max_eta;
        hdr->max_time = acc->max_time; hdr->nbins = 0; hdr->n_events = (int64_t)acc->n_ev;
        hdr->n_event_times = (int64_t)acc->n_times; hdr->pll = pll; hdr->min_log_hz = 0.f; hdr->reserved = 0;
        out_loss[0] = (float)loss;
    }
}

struct SortedLayout {
    size_t off_acc, off_keys, off_vals, off_keys_s, off_idx_s, off_w, off_gs, off_ge, off_D, off_prefE, off_cntE, off_a,
        off_f, off_PA, off_PF, off_tmp, total;
};

SortedLayout sorted_layout(int64_t n) {
    SortedLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.off_acc = take(sizeof(Acc));
    L.off_keys = take(N * 4); L.off_vals = take(N * 4); L.off_keys_s = take(N * 4); L.off_idx_s = take(N * 4);
    L.off_w = take(N * 8); L.off_gs = take(N * 4); L.off_ge = take(N * 4); L.off_D = take(N * 8);
    L.off_prefE = take(N * 8); L.off_cntE = take(N * 8);
    L.off_a = take(N * 8); L.off_f = take(N * 8); L.off_PA = take(N * 8); L.off_PF = take(N * 8);
    size_t tmp = sortscan::radix_sort_temp_bytes((int64_t)N), sc = sortscan::scan_state_bytes((int64_t)N);
    L.off_tmp = take(tmp > sc ? tmp : sc);
    L.total = o;
    return L;
}

}  // namespace

size_t cox_sorted_workspace_bytes(int64_t n) { return sorted_layout(n).total; }

int32_t cox_sorted_fwd_launch(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                              int ties, int reduction, float *out_loss, void *state, size_t state_bytes,
                              void *ws, size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(n >= 1 && n < (int64_t)INT_MAX, "n must be in [1, 2^31)");
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    const SortedLayout L = sorted_layout(n);
    if (ws_bytes < L.total) { set_error("cox sorted: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    const size_t need = sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
    if (state_bytes < need) { set_error("cox sorted: state buffer %zu < %zu", state_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    Acc *acc = reinterpret_cast<Acc *>(w8 + L.off_acc);
    uint32_t *keys = reinterpret_cast<uint32_t *>(w8 + L.off_keys), *vals = reinterpret_cast<uint32_t *>(w8 + L.off_vals);
    uint32_t *keys_s = reinterpret_cast<uint32_t *>(w8 + L.off_keys_s), *idx_s = reinterpret_cast<uint32_t *>(w8 + L.off_idx_s);
    double *w = reinterpret_cast<double *>(w8 + L.off_w);
    int *gs = reinterpret_cast<int *>(w8 + L.off_gs), *ge = reinterpret_cast<int *>(w8 + L.off_ge);
    double *Dpos = reinterpret_cast<double *>(w8 + L.off_D), *prefE = reinterpret_cast<double *>(w8 + L.off_prefE),
           *cntE = reinterpret_cast<double *>(w8 + L.off_cntE), *a = reinterpret_cast<double *>(w8 + L.off_a),
           *f = reinterpret_cast<double *>(w8 + L.off_f), *PA = reinterpret_cast<double *>(w8 + L.off_PA),
           *PF = reinterpret_cast<double *>(w8 + L.off_PF);
    void *tmp = w8 + L.off_tmp;
    int grid = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    if (grid > cap) grid = cap;

    b200surv_cox_header *hdr = static_cast<b200surv_cox_header *>(state);
    float *grad_unit = reinterpret_cast<float *>(hdr + 1);
    int32_t rc;

    k_init_acc<<<1, 32, 0, st>>>(acc);
    // keys are generated into (keys_s, idx_s): four ping-pong passes leave the sorted pairs there
    k_make_keys<<<grid, 256, 0, st>>>(log_hz, time, event, n, keys_s, idx_s, acc);
    rc = sortscan::radix_sort_pairs(keys_s, idx_s, keys, vals, n, 32, tmp, st);
    if (rc) return rc;
    k_gather<<<grid, 256, 0, st>>>(log_hz, idx_s, n, acc, w);
    rc = sortscan::scan_lookback<sortscan::I_MIN, true>(n, LoadR1{w, keys_s, n}, StoreR1{Dpos, ge}, tmp, st);
    if (rc) return rc;
    rc = sortscan::scan_lookback<sortscan::I_MAX, false, true, true>(n, LoadF1{w, keys_s}, StoreF1{prefE, cntE, gs}, tmp, st);
    if (rc) return rc;
    k_terms<<<grid, 256, 0, st>>>(log_hz, keys_s, idx_s, n, ties, Dpos, prefE, cntE, gs, ge, acc, a, f);
    rc = sortscan::scan_lookback<sortscan::I_MAX, false, false, true>(n, LoadF2{a, f, gs}, StoreF2{PA, PF}, tmp, st);
    if (rc) return rc;
    k_grad<<<grid, 256, 0, st>>>(keys_s, idx_s, n, ties, reduction, w, PA, PF, gs, ge, acc, grad_unit, out_loss, hdr);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // namespace b200surv
