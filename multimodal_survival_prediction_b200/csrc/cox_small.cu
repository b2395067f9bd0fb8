// Cox negative partial log-likelihood, SMALL mode: one CTA per cohort of <= 2048 rows.
//
// This is the regime the reference's own training loops live in (n = 2..8 labelled rows per batch,
// scripts/training/partial_modality_training.py:401-410) and the per-fold cohorts of a CV sweep.
// Any non-negative float times.  Everything stays in shared memory:
//   bitonic sort of (time, censored-bit, row) keys -> rows ascending in time, events first in a tie
//   reverse scan  : risk-set sums D                      (fp64)
//   segmented scan: per-tie-group event sums E, counts m (fp64 / int)
//   per event row : Efron/Breslow denominator, its log and reciprocal
//   scans         : P (prefix over all earlier event terms), F (within the tie group)
//   grad_unit[row] = scale * (d - w * (P - d * F)); loss
// Math: oracle/cox.py header.  The gradient for grad_out = 1 is kept in the state buffer; backward
// is a scale by grad_out[seg].
#include "common.cuh"

namespace b200surv {
namespace {

constexpr int SM_THREADS = 1024;
constexpr int SM_MAX = B200SURV_COX_SMALL_MAX;  // 2048

__device__ __forceinline__ uint32_t time_key(float t, bool ev) {
    // non-negative floats order like their bit patterns; LSB = censored, so events sort first
    return (__float_as_uint(t + 0.f) << 1) | (ev ? 0u : 1u);
}

__global__ void __launch_bounds__(SM_THREADS, 1)
cox_small_fwd(const float *__restrict__ log_hz, const float *__restrict__ time,
              const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off, int64_t n_all,
              int ties, int reduction, float *__restrict__ out_loss,
              b200surv_cox_header *__restrict__ hdrs, float *__restrict__ grad_unit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);  // [SM_MAX]
    double *wv = reinterpret_cast<double *>(keys + SM_MAX);                       // weights
    double *sA = wv + SM_MAX;   // phase A: suffix sums of w      | phase B: prefix of a
    double *sB = sA + SM_MAX;   // phase A: segmented sums of w*d | phase B: segmented sums of f
    int *cm = reinterpret_cast<int *>(sB + SM_MAX);                               // segmented counts of d
    unsigned char *fl = reinterpret_cast<unsigned char *>(cm + SM_MAX);           // segment flags
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    __shared__ long long red_l[32];

    const int seg = blockIdx.x, tid = threadIdx.x;
    const int64_t a = seg_off ? seg_off[seg] : 0;
    const int64_t b = seg_off ? seg_off[seg + 1] : n_all;
    const int n = (int)(b - a);
    b200surv_cox_header *hdr = hdrs + seg;
    if (n <= 0 || n > SM_MAX) {  // host validates; defensive
        if (tid == 0) {
            hdr->flags = n > SM_MAX ? B200SURV_COXF_NOT_BINNABLE : 0; hdr->mode = B200SURV_COX_SMALL;
            hdr->loss = n > SM_MAX ? __int_as_float(0x7fc00000) : 0.f; hdr->scale = 0.f; hdr->shift = 0.f;
            hdr->max_log_hz = 0.f; hdr->max_time = 0.f; hdr->nbins = 0; hdr->n_events = 0;
            hdr->n_event_times = 0; hdr->pll = 0.0; hdr->min_log_hz = 0.f; hdr->reserved = 0; out_loss[seg] = hdr->loss;
        }
        return;
    }
    int np = 2;
    while (np < n) np <<= 1;

    // ---- load keys, max log_hz, validity
    float mx = -INFINITY, mt = -INFINITY;
    unsigned flags = 0;
    for (int p = tid; p < np; p += SM_THREADS) {
        unsigned long long k = ~0ull;
        if (p < n) {
            const float t = time[a + p];
            const bool ev = event[a + p] != 0;
            if (!(t >= 0.f)) flags |= B200SURV_COXF_BAD_TIME;
            mx = fmaxf(mx, log_hz[a + p]);
            mt = fmaxf(mt, t);
            k = ((unsigned long long)time_key(t, ev) << 32) | (unsigned)p;
        }
        keys[p] = k;
    }
    mx = block_reduce<float>(mx, -INFINITY, OpMaxF(), red_f);
    mt = block_reduce<float>(mt, -INFINITY, OpMaxF(), red_f);
    flags = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
    const double c = (double)mx;

    // ---- bitonic sort (ascending)
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int p = tid; p < np; p += SM_THREADS) {
                const int q = p ^ j;
                if (q > p) {
                    const unsigned long long x = keys[p], y = keys[q];
                    const bool up = (p & k) == 0;
                    if ((x > y) == up) { keys[p] = y; keys[q] = x; }
                }
            }
        }
    }
    __syncthreads();

    // ---- weights, event flags, tie-group heads
    for (int p = tid; p < np; p += SM_THREADS) {
        double w = 0.0;
        int d = 0;
        unsigned char head = 1;
        if (p < n) {
            const unsigned long long k = keys[p];
            const unsigned tk = (unsigned)(k >> 32);
            d = (tk & 1u) ? 0 : 1;
            w = exp((double)log_hz[a + (unsigned)k] - c);
            head = (p == 0) || ((unsigned)(keys[p - 1] >> 33) != (tk >> 1));
        }
        wv[p] = w; sA[p] = w; sB[p] = d ? w : 0.0; cm[p] = d; fl[p] = head;
    }
    __syncthreads();
    // ---- phase A scans: sA reverse plain; (sB, cm) forward segmented by fl
    for (int dd = 1; dd < np; dd <<= 1) {
        double addA[2], addB[2];
        int addC[2];
        unsigned char nf[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = tid + u * SM_THREADS;
            addA[u] = 0.0; addB[u] = 0.0; addC[u] = 0; nf[u] = 1;
            if (p < np) {
                if (p + dd < np) addA[u] = sA[p + dd];
                nf[u] = fl[p];
                if (p - dd >= 0) {
                    if (!fl[p]) { addB[u] = sB[p - dd]; addC[u] = cm[p - dd]; nf[u] = fl[p - dd]; }
                } else {
                    nf[u] = 1;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = tid + u * SM_THREADS;
            if (p < np) { sA[p] += addA[u]; sB[p] += addB[u]; cm[p] += addC[u]; fl[p] = nf[u]; }
        }
        __syncthreads();
    }
    // ---- per-row group bounds and Efron/Breslow terms (two rows per thread, kept in registers)
    int gs[2], ge[2], dl[2];
    double aterm[2], fterm[2], wrow[2];
    double sum_eta = 0.0, sum_log = 0.0;
    long long n_ev = 0, n_times = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int p = tid + u * SM_THREADS;
        gs[u] = ge[u] = 0; dl[u] = 0; aterm[u] = fterm[u] = 0.0; wrow[u] = 0.0;
        if (p < n) {
            const unsigned long long k = keys[p];
            const unsigned tg = (unsigned)(k >> 33);
            int lo = 0, hi = p;  // first q with group(q) >= tg
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((unsigned)(keys[mid] >> 33) < tg) lo = mid + 1; else hi = mid; }
            gs[u] = lo;
            lo = p + 1; hi = n;  // first q with group(q) > tg
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((unsigned)(keys[mid] >> 33) > tg) hi = mid; else lo = mid + 1; }
            ge[u] = lo;
            dl[u] = (((unsigned)(k >> 32)) & 1u) ? 0 : 1;
            wrow[u] = wv[p];
            if (dl[u]) {
                const double D = sA[gs[u]], E = sB[ge[u] - 1];
                const int m = cm[ge[u] - 1], l = p - gs[u];
                double den = D, frac = 0.0;
                if (ties == B200SURV_TIES_EFRON) { frac = (double)l / (double)m; den = D - frac * E; }
                aterm[u] = 1.0 / den;
                fterm[u] = frac / den;
                sum_log += log(den) + c;
                sum_eta += (double)log_hz[a + (unsigned)k];
                n_ev += 1;
                n_times += (l == 0);
            }
        }
    }
    __syncthreads();  // everyone has read sA/sB/cm -> reuse them for phase B
    for (int u = 0; u < 2; ++u) {
        const int p = tid + u * SM_THREADS;
        if (p < np) {
            sA[p] = aterm[u]; sB[p] = fterm[u];
            fl[p] = (p < n) ? (unsigned char)(p == gs[u]) : 1;
        }
    }
    __syncthreads();
    // ---- phase B scans: sA forward plain; sB forward segmented
    for (int dd = 1; dd < np; dd <<= 1) {
        double addA[2], addB[2];
        unsigned char nf[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = tid + u * SM_THREADS;
            addA[u] = 0.0; addB[u] = 0.0; nf[u] = 1;
            if (p < np) {
                nf[u] = fl[p];
                if (p - dd >= 0) {
                    addA[u] = sA[p - dd];
                    if (!fl[p]) { addB[u] = sB[p - dd]; nf[u] = fl[p - dd]; }
                } else {
                    nf[u] = 1;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = tid + u * SM_THREADS;
            if (p < np) { sA[p] += addA[u]; sB[p] += addB[u]; fl[p] = nf[u]; }
        }
        __syncthreads();
    }
    // ---- reductions, loss, gradient
    sum_eta = block_reduce<double>(sum_eta, 0.0, OpAddD(), red_d);
    sum_log = block_reduce<double>(sum_log, 0.0, OpAddD(), red_d);
    n_ev = block_reduce<long long>(n_ev, 0ll, OpAddLL(), red_l);
    n_times = block_reduce<long long>(n_times, 0ll, OpAddLL(), red_l);
    const double pll = sum_eta - sum_log;
    double norm = 1.0;
    if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)n_ev;
    else if (reduction == B200SURV_REDUCE_MEAN_TERMS)
        norm = (ties == B200SURV_TIES_EFRON) ? (double)n_times : (double)n_ev;
    double scale = (n_ev > 0) ? -1.0 / norm : 0.0;
    double loss = (n_ev > 0) ? -pll / norm : 0.0;
    if (flags) { loss = __longlong_as_double(0x7ff8000000000000ll); scale = loss; }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int p = tid + u * SM_THREADS;
        if (p < n) {
            const double P = sA[ge[u] - 1], F = sB[ge[u] - 1];
            const double d = (double)dl[u];
            const double g = d - wrow[u] * (P - d * F);
            grad_unit[a + (unsigned)keys[p]] = (float)(scale * g);
        }
    }
    if (tid == 0) {
        hdr->flags = flags; hdr->mode = B200SURV_COX_SMALL; hdr->loss = (float)loss; hdr->scale = (float)scale;
        hdr->shift = mx; hdr->max_log_hz = mx; hdr->max_time = mt; hdr->nbins = 0; hdr->n_events = n_ev;
        hdr->n_event_times = n_times; hdr->pll = pll; hdr->min_log_hz = 0.f; hdr->reserved = 0;
        out_loss[seg] = (float)loss;
    }
}

}  // namespace

// out[i] = grad_out[seg(i)] * grad_unit[i]   (SMALL and SORTED modes keep the unit gradient)
__global__ void __launch_bounds__(256)
cox_scale_grad(const float *__restrict__ grad_out, const float *__restrict__ grad_unit,
               const int64_t *__restrict__ seg_off, int64_t n, float *__restrict__ out) {
    const int seg = blockIdx.y;
    const int64_t a = seg_off ? seg_off[seg] : 0, b = seg_off ? seg_off[seg + 1] : n;
    const float g = grad_out[seg];
    for (int64_t i = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = g * grad_unit[i];
}

size_t cox_small_smem_bytes() {
    return (size_t)SM_MAX * (sizeof(unsigned long long) + 3 * sizeof(double) + sizeof(int) + 1);
}

int32_t cox_small_fwd_launch(const float *log_hz, const float *time, const uint8_t *event,
                             const int64_t *seg_off, int64_t n, int64_t n_seg, int ties, int reduction,
                             float *out_loss, void *state, size_t state_bytes, cudaStream_t st) {
    B200_REQUIRE(n_seg >= 1, "n_seg");
    B200_REQUIRE(seg_off != nullptr || n <= SM_MAX, "SMALL mode needs every segment <= 2048 rows");
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    const size_t need = (size_t)n_seg * sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
    if (state_bytes < need) { set_error("cox small: state buffer %zu < %zu", state_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_small_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)cox_small_smem_bytes()));
        attr_once.mark();
    }
    b200surv_cox_header *hdrs = static_cast<b200surv_cox_header *>(state);
    float *grad_unit = reinterpret_cast<float *>(hdrs + n_seg);
    cox_small_fwd<<<(unsigned)n_seg, SM_THREADS, cox_small_smem_bytes(), st>>>(
        log_hz, time, event, seg_off, n, ties, reduction, out_loss, hdrs, grad_unit);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(1);
    return B200SURV_OK;
}

int32_t cox_scale_grad_launch(const float *grad_out, const void *state, const int64_t *seg_off, int64_t n,
                              int64_t n_seg, float *out_grad, cudaStream_t st) {
    if (n == 0) return B200SURV_OK;
    const b200surv_cox_header *hdrs = static_cast<const b200surv_cox_header *>(state);
    const float *grad_unit = reinterpret_cast<const float *>(hdrs + n_seg);
    int64_t per = n / n_seg + 1;
    int gx = (int)((per + 255) / 256);
    const int cap = n_seg == 1 ? 8 * num_sms() : (int)((8 * num_sms() + n_seg - 1) / n_seg);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    cox_scale_grad<<<dim3(gx, (unsigned)n_seg), 256, 0, st>>>(grad_out, grad_unit, seg_off, n, out_grad);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(1);
    return B200SURV_OK;
}

}  // namespace b200surv
