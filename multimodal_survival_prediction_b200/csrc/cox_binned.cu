// Cox negative partial log-likelihood, BINNED mode: integer-valued times in [0, nbins).
//
// Replaces torchsurv.loss.cox.neg_partial_log_likelihood as called by the reference at
// scripts/training/partial_modality_training.py:285-288 (math: oracle/cox.py header).
//
// Breslow/Efron depend on time only through per-distinct-time aggregates, so no sort is needed:
//   pass 1  (stream log_hz,time,event once: 9 B/row)  per-CTA shared-memory histograms
//           S_cens[b] += w, (S_event[b], m[b]) += (w, 1),  w = exp(log_hz - shift)
//   reduce  per-CTA partials -> per-bin sums in fp64 (deterministic order)
//   scan    D[b] = sum_{b' >= b} S[b']  (risk-set sums, fp64), event offsets, log D, E/D
//   items   Efron: one work item per event (b, l<m_b): log/recip of 1 - (l/m) E/D
//   finish  P[b] = sum_{b' <= b} G[b'], F[b]; loss; (P,F) table kept for backward
//   pass 2  (stream the rows again: 9 B read + 4 B write)  grad = scale*(d - w*(P[b] - d*F[b]))
// Algorithmic HBM bytes: 22 per row for fwd+bwd; everything else is O(nbins).
#include <climits>

#include "common.cuh"

namespace b200surv {
namespace {

constexpr int P1_THREADS = 1024;
constexpr int P2_THREADS = 512;
constexpr int FIN_THREADS = 1024;
constexpr float EXP_RANGE_LIMIT = 60.f;

struct CtaRec {
    double sum_ev_eta;
    float max_eta;
    float max_time;
    unsigned flags;
    unsigned pad;
};

struct SegRange {
    int64_t a, b;    // rows [a, b)
    int64_t va, vb;  // 4-aligned interior [va, vb), processed as 128-bit groups
};
__device__ __forceinline__ SegRange seg_range(const int64_t *seg_off, int64_t n, int seg, bool vec_ok) {
    SegRange r;
    r.a = seg_off ? seg_off[seg] : 0;
    r.b = seg_off ? seg_off[seg + 1] : n;
    if (vec_ok) {
        r.va = min(r.b, (r.a + 3) & ~int64_t(3));
        r.vb = r.va + ((r.b - r.va) & ~int64_t(3));
    } else {
        r.va = r.vb = r.a;  // everything through the scalar path
    }
    return r;
}

// ---------------------------------------------------------------- pass 1: accumulate
struct P1Acc {
    float mx, mt, se;
    unsigned flags;
};

__device__ __forceinline__ void p1_row(float eta, float t, unsigned ev, float shift, float nbf,
                                       unsigned long long *h_ev, float *h_c, P1Acc &acc) {
    const float w = __expf(eta - shift);
    acc.mx = fmaxf(acc.mx, eta);
    acc.mt = fmaxf(acc.mt, t);
    const bool ok = (t >= 0.f) && (t < nbf) && (t == truncf(t));
    if (ok) {
        const int bin = (int)t;
        if (ev) {
            acc.se += eta;
            // one 64-bit CAS updates (sum, count) of the bin's event rows together
            unsigned long long *p = h_ev + bin;
            unsigned long long old = *p, assumed;
            do {
                assumed = old;
                const float s = __uint_as_float((unsigned)assumed) + w;
                const unsigned long long m = (assumed >> 32) + 1ull;
                old = atomicCAS(p, assumed, (m << 32) | (unsigned long long)__float_as_uint(s));
            } while (old != assumed);
        } else {
            atomicAdd(h_c + bin, w);
        }
    } else {
        acc.flags |= (t >= 0.f) ? B200SURV_COXF_NOT_BINNABLE
                                : (B200SURV_COXF_NOT_BINNABLE | B200SURV_COXF_BAD_TIME);
        if (ev) acc.se += eta;
    }
}

// partial layout per (seg, cta): float S_cens[nb], float S_event[nb], uint32 m[nb]
__global__ void __launch_bounds__(P1_THREADS, 1)
cox_binned_pass1(const float *__restrict__ log_hz, const float *__restrict__ time,
                 const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off, int64_t n,
                 int nb, float shift, int vec_ok, float *__restrict__ partial,
                 CtaRec *__restrict__ recs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *h_ev = reinterpret_cast<unsigned long long *>(smem_raw);
    float *h_c = reinterpret_cast<float *>(smem_raw + sizeof(unsigned long long) * nb);
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];

    const int seg = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) { h_ev[i] = 0ull; h_c[i] = 0.f; }
    __syncthreads();

    const SegRange r = seg_range(seg_off, n, seg, vec_ok != 0);
    const float nbf = (float)nb;
    P1Acc acc{-INFINITY, -INFINITY, 0.f, 0u};

    // 128-bit groups, two per thread per iteration (all six loads issued before any use)
    const int64_t ngroups = (r.vb - r.va) >> 2;
    const int64_t stride = (int64_t)nctas * blockDim.x;
    int64_t g = (int64_t)cta * blockDim.x + threadIdx.x;
    const float *lh = log_hz + r.va;
    const float *tm = time + r.va;
    const uint8_t *evp = event + r.va;
    for (; g + stride < ngroups; g += 2 * stride) {
        const int64_t g2 = g + stride;
        const float4 e0 = ldg_stream_f4(lh + 4 * g), e1 = ldg_stream_f4(lh + 4 * g2);
        const float4 t0 = ldg_stream_f4(tm + 4 * g), t1 = ldg_stream_f4(tm + 4 * g2);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g), v1 = ldg_stream_u32(evp + 4 * g2);
        p1_row(e0.x, t0.x, v0 & 0xffu, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.y, t0.y, v0 & 0xff00u, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.z, t0.z, v0 & 0xff0000u, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.w, t0.w, v0 & 0xff000000u, shift, nbf, h_ev, h_c, acc);
        p1_row(e1.x, t1.x, v1 & 0xffu, shift, nbf, h_ev, h_c, acc);
        p1_row(e1.y, t1.y, v1 & 0xff00u, shift, nbf, h_ev, h_c, acc);
        p1_row(e1.z, t1.z, v1 & 0xff0000u, shift, nbf, h_ev, h_c, acc);
        p1_row(e1.w, t1.w, v1 & 0xff000000u, shift, nbf, h_ev, h_c, acc);
    }
    if (g < ngroups) {
        const float4 e0 = ldg_stream_f4(lh + 4 * g);
        const float4 t0 = ldg_stream_f4(tm + 4 * g);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g);
        p1_row(e0.x, t0.x, v0 & 0xffu, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.y, t0.y, v0 & 0xff00u, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.z, t0.z, v0 & 0xff0000u, shift, nbf, h_ev, h_c, acc);
        p1_row(e0.w, t0.w, v0 & 0xff000000u, shift, nbf, h_ev, h_c, acc);
    }
    // unaligned head [a, va) and tail [vb, b): scalar, spread over the CTAs
    {
        const int64_t nhead = r.va - r.a, ntail = r.b - r.vb;
        for (int64_t k = (int64_t)cta * blockDim.x + threadIdx.x; k < nhead + ntail; k += stride) {
            const int64_t row = (k < nhead) ? (r.a + k) : (r.vb + (k - nhead));
            p1_row(log_hz[row], time[row], event[row], shift, nbf, h_ev, h_c, acc);
        }
    }
    __syncthreads();

    // flush the CTA histogram with plain coalesced stores (reduced in fp64 by the next kernel)
    float *out = partial + ((size_t)seg * nctas + cta) * 3 * (size_t)nb;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const unsigned long long v = h_ev[i];
        out[i] = h_c[i];
        out[nb + i] = __uint_as_float((unsigned)v);
        reinterpret_cast<unsigned *>(out)[2 * nb + i] = (unsigned)(v >> 32);
    }
    const double se = block_reduce<double>((double)acc.se, 0.0, OpAddD(), red_d);
    const float mx = block_reduce<float>(acc.mx, -INFINITY, OpMaxF(), red_f);
    const float mt = block_reduce<float>(acc.mt, -INFINITY, OpMaxF(), red_f);
    const unsigned fl = block_reduce<unsigned>(acc.flags, 0u, OpOrU(), red_u);
    if (threadIdx.x == 0) {
        CtaRec rec;
        rec.sum_ev_eta = se; rec.max_eta = mx; rec.max_time = mt; rec.flags = fl; rec.pad = 0;
        recs[(size_t)seg * nctas + cta] = rec;
    }
}

// ---------------------------------------------------------------- reduce partials -> bins (fp64)
// bins_sum layout per segment: S_all[nb], S_event[nb], m[nb], sum_ev_eta, n_not_binnable,
// n_exp_range(unused here), n_bad_time.   bins_max: max_eta, max_time.
__global__ void __launch_bounds__(256)
cox_binned_reduce(const float *__restrict__ partial, const CtaRec *__restrict__ recs, int nctas,
                  int nb, double *__restrict__ bins_sum, float *__restrict__ bins_max) {
    const int seg = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cnt = 3 * (size_t)nb + 4;
    double *bs = bins_sum + (size_t)seg * cnt;
    if (b < nb) {
        double sc = 0.0, se = 0.0;
        unsigned long long m = 0;
        const float *p = partial + (size_t)seg * nctas * 3 * (size_t)nb;
        for (int c = 0; c < nctas; ++c, p += 3 * (size_t)nb) {
            sc += (double)p[b];
            se += (double)p[nb + b];
            m += reinterpret_cast<const unsigned *>(p)[2 * nb + b];
        }
        bs[b] = sc + se;
        bs[nb + b] = se;
        bs[2 * nb + b] = (double)m;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double se = 0.0;
        float mx = -INFINITY, mt = -INFINITY;
        unsigned fl = 0;
        for (int c = threadIdx.x; c < nctas; c += 32) {
            const CtaRec r = recs[(size_t)seg * nctas + c];
            se += r.sum_ev_eta; mx = fmaxf(mx, r.max_eta); mt = fmaxf(mt, r.max_time); fl |= r.flags;
        }
        se = warp_sum(se); mx = warp_max(mx); mt = warp_max(mt); fl = warp_or(fl);
        if (threadIdx.x == 0) {
            bs[3 * (size_t)nb + 0] = se;
            bs[3 * (size_t)nb + 1] = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1.0 : 0.0;
            bs[3 * (size_t)nb + 2] = 0.0;
            bs[3 * (size_t)nb + 3] = (fl & B200SURV_COXF_BAD_TIME) ? 1.0 : 0.0;
            bins_max[2 * seg + 0] = mx;
            bins_max[2 * seg + 1] = mt;
        }
    }
}

// ---------------------------------------------------------------- block scans (one CTA of 1024)
// exclusive scan of one double per thread; returns the exclusive prefix, *total = block total
__device__ __forceinline__ double block_exscan_d(double v, double *sh /*[32]*/, double *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        double w = (lane < nw) ? sh[lane] : 0.0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += u;
        }
        sh[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    *total = sh[32];
    return sh[wid] + (inc - v);
}

struct FinScratch {  // per-segment scratch in the workspace
    // double D[nb], double rED[nb] (E/D), double logD[nb], int32 moff[nb+1], double TGF[3][nb]
};

// scan: D (suffix), moff (exclusive prefix of m), logD, E/D; zero TGF accumulators; totals
__global__ void __launch_bounds__(FIN_THREADS, 1)
cox_binned_scan(const double *__restrict__ bins_sum, int nb, double *__restrict__ Dg,
                double *__restrict__ rg, double *__restrict__ logDg, int *__restrict__ moff,
                double *__restrict__ tgf, long long *__restrict__ totals) {
    __shared__ double sh[33];
    const int seg = blockIdx.x;
    const size_t cnt = 3 * (size_t)nb + 4;
    const double *bs = bins_sum + (size_t)seg * cnt;
    Dg += (size_t)seg * nb; rg += (size_t)seg * nb; logDg += (size_t)seg * nb;
    moff += (size_t)seg * (nb + 1); tgf += (size_t)seg * 3 * nb;
    const int per = (nb + FIN_THREADS - 1) / FIN_THREADS;  // <= 16
    const int t = threadIdx.x;

    // suffix sums of S: thread t owns the reversed chunk, i.e. bins [hi-per, hi)
    {
        const int hi = nb - t * per, lo = max(0, hi - per);
        double loc = 0.0;
        for (int b = hi - 1; b >= lo && b >= 0; --b) loc += bs[b];
        double tot;
        double run = block_exscan_d(hi > 0 ? loc : 0.0, sh, &tot);  // sum of all later chunks
        for (int b = hi - 1; b >= lo && b >= 0; --b) {
            run += bs[b];
            const double D = run, E = bs[nb + b];
            Dg[b] = D;
            const bool has = bs[2 * nb + b] > 0.0;
            rg[b] = has ? E / D : 0.0;
            logDg[b] = has ? log(D) : 0.0;
        }
    }
    // exclusive prefix of m (event offsets), number of event times
    {
        const int lo = t * per, hi = min(nb, lo + per);
        double loc = 0.0;
        long long net = 0;
        for (int b = lo; b < hi; ++b) { const double m = bs[2 * nb + b]; loc += m; net += (m > 0.0); }
        double tot;
        double run = block_exscan_d(loc, sh, &tot);
        for (int b = lo; b < hi; ++b) { moff[b] = (int)run; run += bs[2 * nb + b]; }
        if (t == 0) moff[nb] = (int)tot;
        __shared__ long long shl[32];
        const long long nets = block_reduce<long long>(net, 0ll, OpAddLL(), shl);
        if (t == 0) { totals[2 * seg + 0] = (long long)tot; totals[2 * seg + 1] = nets; }
    }
    for (int i = t; i < 3 * nb; i += blockDim.x) tgf[i] = 0.0;
}

// Efron work items: item w in [0, n_events) -> (bin b, l = w - moff[b]); x = 1 - (l/m) E/D
__global__ void __launch_bounds__(256)
cox_binned_efron_items(int nb, const int *__restrict__ moff_g, const double *__restrict__ rg,
                       double *__restrict__ tgf_g) {
    extern __shared__ int s_moff[];  // nb + 1
    const int seg = blockIdx.y;
    const int *moff = moff_g + (size_t)seg * (nb + 1);
    const double *r = rg + (size_t)seg * nb;
    double *tgf = tgf_g + (size_t)seg * 3 * nb;
    for (int i = threadIdx.x; i <= nb; i += blockDim.x) s_moff[i] = moff[i];
    __syncthreads();
    const int total = s_moff[nb];
    const int lane = threadIdx.x & 31;
    // whole warps iterate together so that the segmented shuffles stay convergent
    for (long long base = (long long)(blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total;
         base += (long long)gridDim.x * blockDim.x) {
        const int w = (int)min(base + lane, (long long)INT_MAX);
        const bool live = w < total;
        int b = nb;  // sentinel bin for idle lanes (never written)
        float vt = 0.f, vg = 0.f, vf = 0.f;
        if (live) {
            int lo = 0, hi = nb - 1;  // first b with moff[b+1] > w
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_moff[mid + 1] > w) hi = mid; else lo = mid + 1;
            }
            b = lo;
            const int m0 = s_moff[b], m = s_moff[b + 1] - m0, l = w - m0;
            const double frac = (double)l / (double)m;
            const float x = (float)(1.0 - frac * r[b]);
            const float rx = __frcp_rn(x);
            vt = __logf(x);
            vg = rx;
            vf = (float)frac * rx;
        }
        // segmented (by bin) inclusive suffix-reduction across the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int bo = __shfl_down_sync(FULL, b, o);
            const float t1 = __shfl_down_sync(FULL, vt, o), g1 = __shfl_down_sync(FULL, vg, o),
                        f1 = __shfl_down_sync(FULL, vf, o);
            if (lane + o < 32 && bo == b) { vt += t1; vg += g1; vf += f1; }
        }
        const int bprev = __shfl_up_sync(FULL, b, 1);
        if (live && (lane == 0 || bprev != b)) {
            atomicAdd(tgf + b, (double)vt);
            atomicAdd(tgf + nb + b, (double)vg);
            atomicAdd(tgf + 2 * nb + b, (double)vf);
        }
    }
}

// finish: T,G,F per bin -> P prefix, loss, header, (P,F) table
__global__ void __launch_bounds__(FIN_THREADS, 1)
cox_binned_finish(const double *__restrict__ bins_sum, const float *__restrict__ bins_max, int nb,
                  int ties, int reduction, float shift, const double *__restrict__ Dg,
                  const double *__restrict__ logDg, const double *__restrict__ tgf_g,
                  const long long *__restrict__ totals, float *__restrict__ out_loss,
                  unsigned char *__restrict__ state, size_t seg_state_stride) {
    __shared__ double sh[33];
    __shared__ double shd[32];
    const int seg = blockIdx.x, t = threadIdx.x;
    const size_t cnt = 3 * (size_t)nb + 4;
    const double *bs = bins_sum + (size_t)seg * cnt;
    const double *D = Dg + (size_t)seg * nb, *logD = logDg + (size_t)seg * nb;
    const double *tgf = tgf_g + (size_t)seg * 3 * nb;
    b200surv_cox_header *hdr = reinterpret_cast<b200surv_cox_header *>(state + seg * seg_state_stride);
    float2 *table = reinterpret_cast<float2 *>(state + seg * seg_state_stride + sizeof(b200surv_cox_header));

    const int per = (nb + FIN_THREADS - 1) / FIN_THREADS;
    const int lo = t * per, hi = min(nb, lo + per);
    double gl[16], fl[16];
    double gsum = 0.0, tsum = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int b = lo + k;
        double G = 0.0, F = 0.0;
        if (k < per && b < hi) {
            const double m = bs[2 * nb + b];
            if (m > 0.0) {
                const double invD = 1.0 / D[b];
                if (ties == B200SURV_TIES_EFRON) {
                    tsum += m * logD[b] + tgf[b];
                    G = tgf[nb + b] * invD;
                    F = tgf[2 * nb + b] * invD;
                } else {
                    tsum += m * logD[b];
                    G = m * invD;
                }
            }
        }
        gl[k] = G; fl[k] = F; gsum += G;
    }
    double tot;
    double run = block_exscan_d(gsum, sh, &tot);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int b = lo + k;
        if (k < per && b < hi) {
            run += gl[k];
            table[b] = make_float2((float)run, (float)fl[k]);
        }
    }
    const double T = block_reduce<double>(tsum, 0.0, OpAddD(), shd);
    if (t == 0) {
        const long long n_events = totals[2 * seg + 0], n_times = totals[2 * seg + 1];
        const double pll = bs[3 * (size_t)nb] - (T + (double)n_events * (double)shift);
        double norm = 1.0;
        if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)n_events;
        else if (reduction == B200SURV_REDUCE_MEAN_TERMS)
            norm = (ties == B200SURV_TIES_EFRON) ? (double)n_times : (double)n_events;
        unsigned flags = 0;
        if (bs[3 * (size_t)nb + 1] != 0.0) flags |= B200SURV_COXF_NOT_BINNABLE;
        if (bs[3 * (size_t)nb + 3] != 0.0) flags |= B200SURV_COXF_BAD_TIME;
        const float mx = bins_max[2 * seg];
        if (mx - shift > EXP_RANGE_LIMIT) flags |= B200SURV_COXF_EXP_RANGE;
        float loss = 0.f, scale = 0.f;
        if (n_events > 0) { loss = (float)(-pll / norm); scale = (float)(-1.0 / norm); }
        if (flags) { loss = __int_as_float(0x7fc00000); scale = loss; }
        hdr->flags = flags; hdr->mode = B200SURV_COX_BINNED; hdr->loss = loss; hdr->scale = scale;
        hdr->shift = shift; hdr->max_log_hz = mx; hdr->max_time = bins_max[2 * seg + 1];
        hdr->nbins = nb; hdr->n_events = n_events; hdr->n_event_times = n_times; hdr->pll = pll;
        hdr->reserved = 0;
        out_loss[seg] = loss;
    }
}

// ---------------------------------------------------------------- pass 2: gradient
__device__ __forceinline__ float p2_row(float eta, float t, unsigned ev, float shift, float k,
                                        const float2 *tab, int nb) {
    const float w = __expf(eta - shift);
    int bin = (int)t;
    bin = min(max(bin, 0), nb - 1);  // invalid rows are already poisoned through k = NaN
    const float2 pf = tab[bin];
    const float d = ev ? 1.f : 0.f;
    return k * (d - w * (pf.x - d * pf.y));
}

__global__ void __launch_bounds__(P2_THREADS, 2)
cox_binned_bwd(const float *__restrict__ grad_out, const unsigned char *__restrict__ state,
               size_t seg_state_stride, const float *__restrict__ log_hz,
               const float *__restrict__ time, const uint8_t *__restrict__ event,
               const int64_t *__restrict__ seg_off, int64_t n, int nb, int vec_ok,
               float *__restrict__ out_grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tab = reinterpret_cast<float2 *>(smem_raw);
    const int seg = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x;
    const b200surv_cox_header *hdr =
        reinterpret_cast<const b200surv_cox_header *>(state + seg * seg_state_stride);
    const float2 *gtab =
        reinterpret_cast<const float2 *>(state + seg * seg_state_stride + sizeof(b200surv_cox_header));
    for (int i = threadIdx.x; i < nb; i += blockDim.x) tab[i] = gtab[i];
    const float shift = hdr->shift;
    const float k = hdr->scale * grad_out[seg];
    __syncthreads();

    const SegRange r = seg_range(seg_off, n, seg, vec_ok != 0);
    const int64_t ngroups = (r.vb - r.va) >> 2;
    const int64_t stride = (int64_t)nctas * blockDim.x;
    int64_t g = (int64_t)cta * blockDim.x + threadIdx.x;
    const float *lh = log_hz + r.va;
    const float *tm = time + r.va;
    const uint8_t *evp = event + r.va;
    float *og = out_grad + r.va;
    for (; g + stride < ngroups; g += 2 * stride) {
        const int64_t g2 = g + stride;
        const float4 e0 = ldg_stream_f4(lh + 4 * g), e1 = ldg_stream_f4(lh + 4 * g2);
        const float4 t0 = ldg_stream_f4(tm + 4 * g), t1 = ldg_stream_f4(tm + 4 * g2);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g), v1 = ldg_stream_u32(evp + 4 * g2);
        float4 o0, o1;
        o0.x = p2_row(e0.x, t0.x, v0 & 0xffu, shift, k, tab, nb);
        o0.y = p2_row(e0.y, t0.y, v0 & 0xff00u, shift, k, tab, nb);
        o0.z = p2_row(e0.z, t0.z, v0 & 0xff0000u, shift, k, tab, nb);
        o0.w = p2_row(e0.w, t0.w, v0 & 0xff000000u, shift, k, tab, nb);
        o1.x = p2_row(e1.x, t1.x, v1 & 0xffu, shift, k, tab, nb);
        o1.y = p2_row(e1.y, t1.y, v1 & 0xff00u, shift, k, tab, nb);
        o1.z = p2_row(e1.z, t1.z, v1 & 0xff0000u, shift, k, tab, nb);
        o1.w = p2_row(e1.w, t1.w, v1 & 0xff000000u, shift, k, tab, nb);
        stg_stream_f4(og + 4 * g, o0);
        stg_stream_f4(og + 4 * g2, o1);
    }
    if (g < ngroups) {
        const float4 e0 = ldg_stream_f4(lh + 4 * g);
        const float4 t0 = ldg_stream_f4(tm + 4 * g);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g);
        float4 o0;
        o0.x = p2_row(e0.x, t0.x, v0 & 0xffu, shift, k, tab, nb);
        o0.y = p2_row(e0.y, t0.y, v0 & 0xff00u, shift, k, tab, nb);
        o0.z = p2_row(e0.z, t0.z, v0 & 0xff0000u, shift, k, tab, nb);
        o0.w = p2_row(e0.w, t0.w, v0 & 0xff000000u, shift, k, tab, nb);
        stg_stream_f4(og + 4 * g, o0);
    }
    {
        const int64_t nhead = r.va - r.a, ntail = r.b - r.vb;
        for (int64_t q = (int64_t)cta * blockDim.x + threadIdx.x; q < nhead + ntail; q += stride) {
            const int64_t row = (q < nhead) ? (r.a + q) : (r.vb + (q - nhead));
            out_grad[row] = p2_row(log_hz[row], time[row], event[row], shift, k, tab, nb);
        }
    }
}

// ---------------------------------------------------------------- host-side layout
struct BinnedLayout {
    int nctas;            // pass-1 CTAs per segment
    size_t off_partial, off_recs, off_bins_sum, off_bins_max, off_D, off_r, off_logD, off_moff,
        off_tgf, off_totals, total;
};

BinnedLayout binned_layout(int64_t n, int64_t n_seg, int nb) {
    BinnedLayout L;
    const int sms = num_sms();
    int64_t c;
    if (n_seg == 1) {
        c = (n + 4 * P1_THREADS - 1) / (4 * P1_THREADS);
        if (c > sms) c = sms;
    } else {
        c = (2 * sms + n_seg - 1) / n_seg;
        const int64_t by_rows = (n / n_seg + 4 * P1_THREADS - 1) / (4 * P1_THREADS);
        if (c > by_rows) c = by_rows;
    }
    if (c < 1) c = 1;
    L.nctas = (int)c;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    L.off_partial = take((size_t)n_seg * L.nctas * 3 * nb * sizeof(float));
    L.off_recs = take((size_t)n_seg * L.nctas * sizeof(CtaRec));
    L.off_bins_sum = take((size_t)n_seg * (3 * (size_t)nb + 4) * sizeof(double));
    L.off_bins_max = take((size_t)n_seg * 2 * sizeof(float));
    L.off_D = take((size_t)n_seg * nb * sizeof(double));
    L.off_r = take((size_t)n_seg * nb * sizeof(double));
    L.off_logD = take((size_t)n_seg * nb * sizeof(double));
    L.off_moff = take((size_t)n_seg * (nb + 1) * sizeof(int));
    L.off_tgf = take((size_t)n_seg * 3 * nb * sizeof(double));
    L.off_totals = take((size_t)n_seg * 2 * sizeof(long long));
    L.total = o;
    return L;
}

inline size_t seg_state_stride(int nb) { return sizeof(b200surv_cox_header) + (size_t)nb * sizeof(float2); }

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int32_t check_nbins(int nb) {
    B200_REQUIRE(nb >= 32 && nb <= B200SURV_COX_MAX_BINS && (nb & (nb - 1)) == 0,
                 "nbins must be a power of two in [32, 16384]");
    return B200SURV_OK;
}

}  // namespace

// ---------------------------------------------------------------- internal entry points
size_t cox_binned_state_bytes(int64_t n_seg, int nb) { return (size_t)n_seg * seg_state_stride(nb); }
size_t cox_binned_workspace_bytes(int64_t n, int64_t n_seg, int nb) { return binned_layout(n, n_seg, nb).total; }

int32_t cox_binned_partial(const float *log_hz, const float *time, const uint8_t *event,
                           const int64_t *seg_off, int64_t n, int64_t n_seg, int nb, float shift,
                           double *bins_sum, float *bins_max, void *ws, size_t ws_bytes,
                           cudaStream_t st) {
    int32_t rc = check_nbins(nb);
    if (rc) return rc;
    B200_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "n must be < 2^31");
    B200_REQUIRE(n_seg >= 1 && n_seg <= 65535, "n_seg must be in [1, 65535]");
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w = static_cast<unsigned char *>(ws);
    float *partial = reinterpret_cast<float *>(w + L.off_partial);
    CtaRec *recs = reinterpret_cast<CtaRec *>(w + L.off_recs);
    const int vec_ok = aligned16(log_hz) && aligned16(time) && ((reinterpret_cast<uintptr_t>(event) & 3) == 0);
    const size_t smem = (size_t)nb * 12;
    static bool attr_done = false;
    if (!attr_done) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             B200SURV_COX_MAX_BINS * 12));
        attr_done = true;
    }
    cox_binned_pass1<<<dim3(L.nctas, (unsigned)n_seg), P1_THREADS, smem, st>>>(
        log_hz, time, event, seg_off, n, nb, shift, vec_ok, partial, recs);
    cox_binned_reduce<<<dim3((nb + 255) / 256, (unsigned)n_seg), 256, 0, st>>>(partial, recs, L.nctas, nb,
                                                                               bins_sum, bins_max);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t cox_binned_finalize(const double *bins_sum, const float *bins_max, int64_t n, int64_t n_seg,
                            int ties, int reduction, int nb, float shift, float *out_loss, void *state,
                            size_t state_bytes, void *ws, size_t ws_bytes, cudaStream_t st) {
    int32_t rc = check_nbins(nb);
    if (rc) return rc;
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    if (state_bytes < cox_binned_state_bytes(n_seg, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w = static_cast<unsigned char *>(ws);
    double *D = reinterpret_cast<double *>(w + L.off_D), *r = reinterpret_cast<double *>(w + L.off_r),
           *logD = reinterpret_cast<double *>(w + L.off_logD), *tgf = reinterpret_cast<double *>(w + L.off_tgf);
    int *moff = reinterpret_cast<int *>(w + L.off_moff);
    long long *totals = reinterpret_cast<long long *>(w + L.off_totals);
    cox_binned_scan<<<(unsigned)n_seg, FIN_THREADS, 0, st>>>(bins_sum, nb, D, r, logD, moff, tgf, totals);
    if (ties == B200SURV_TIES_EFRON) {
        const size_t smem = (size_t)(nb + 1) * sizeof(int);
        static bool attr_done = false;
        if (!attr_done) {
            B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_efron_items, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (B200SURV_COX_MAX_BINS + 1) * (int)sizeof(int)));
            attr_done = true;
        }
        int gx = n_seg == 1 ? 4 * num_sms() : (int)((4 * num_sms() + n_seg - 1) / n_seg);
        if (gx < 1) gx = 1;
        cox_binned_efron_items<<<dim3(gx, (unsigned)n_seg), 256, smem, st>>>(nb, moff, r, tgf);
    }
    cox_binned_finish<<<(unsigned)n_seg, FIN_THREADS, 0, st>>>(bins_sum, bins_max, nb, ties, reduction, shift, D,
                                                               logD, tgf, totals, out_loss,
                                                               static_cast<unsigned char *>(state),
                                                               seg_state_stride(nb));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t cox_binned_fwd(const float *log_hz, const float *time, const uint8_t *event, const int64_t *seg_off,
                       int64_t n, int64_t n_seg, int ties, int reduction, int nb, float shift,
                       float *out_loss, void *state, size_t state_bytes, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
    int32_t rc = check_nbins(nb);
    if (rc) return rc;
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w = static_cast<unsigned char *>(ws);
    double *bins_sum = reinterpret_cast<double *>(w + L.off_bins_sum);
    float *bins_max = reinterpret_cast<float *>(w + L.off_bins_max);
    rc = cox_binned_partial(log_hz, time, event, seg_off, n, n_seg, nb, shift, bins_sum, bins_max, ws, ws_bytes, st);
    if (rc) return rc;
    return cox_binned_finalize(bins_sum, bins_max, n, n_seg, ties, reduction, nb, shift, out_loss, state,
                               state_bytes, ws, ws_bytes, st);
}

int32_t cox_binned_bwd_launch(const float *grad_out, const void *state, size_t state_bytes, const float *log_hz,
                              const float *time, const uint8_t *event, const int64_t *seg_off, int64_t n,
                              int64_t n_seg, int nb, float *out_grad, cudaStream_t st) {
    int32_t rc = check_nbins(nb);
    if (rc) return rc;
    if (state_bytes < cox_binned_state_bytes(n_seg, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    if (n == 0) return B200SURV_OK;
    const int vec_ok = aligned16(log_hz) && aligned16(time) && aligned16(out_grad) &&
                       ((reinterpret_cast<uintptr_t>(event) & 3) == 0);
    const size_t smem = (size_t)nb * sizeof(float2);
    static bool attr_done = false;
    if (!attr_done) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             B200SURV_COX_MAX_BINS * (int)sizeof(float2)));
        attr_done = true;
    }
    const int sms = num_sms();
    int64_t c;
    const int64_t rows_per_cta = 4 * P2_THREADS;
    if (n_seg == 1) { c = (n + rows_per_cta - 1) / rows_per_cta; if (c > 2 * sms) c = 2 * sms; }
    else { c = (4 * sms + n_seg - 1) / n_seg; const int64_t by = (n / n_seg + rows_per_cta - 1) / rows_per_cta; if (c > by) c = by; }
    if (c < 1) c = 1;
    cox_binned_bwd<<<dim3((unsigned)c, (unsigned)n_seg), P2_THREADS, smem, st>>>(
        grad_out, static_cast<const unsigned char *>(state), seg_state_stride(nb), log_hz, time, event, seg_off, n,
        nb, vec_ok, out_grad);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // namespace b200surv
