// Cox negative partial log-likelihood, BINNED mode: integer-valued times in [0, nbins).
//
// Replaces torchsurv.loss.cox.neg_partial_log_likelihood as called by the reference at
// scripts/training/partial_modality_training.py:285-288 (math: oracle/cox.py header).
//
// Breslow/Efron depend on time only through per-distinct-time aggregates, so no sort is needed:
//   K1 pass 1  stream (log_hz,time,event) once, 9 B/row.  Per-CTA shared-memory histograms of
//              w = exp(log_hz - shift) in 32.32 FIXED POINT, accumulated with native 32-bit shared
//              atomics (low word, carry into the high word) + an event counter per bin.  Integer
//              accumulation is exact and associative: the per-bin sums, hence the loss, do not
//              depend on the grid, the order of the atomics or how rows are sharded over GPUs.
//              (fp32 atomicAdd in shared memory compiles to a CAS loop and is 7x slower --
//              measured, profiles/r1_hist_microbench.txt.)
//   K2 reduce  per-CTA partials -> per-bin int64 sums (this is what multi-GPU all-reduces)
//   K3 items   every CTA redundantly suffix-scans the nbins sums in shared memory (fp64):
//              D[b] = sum_{b' >= b} S[b'] and the Efron task offsets; then one warp per
//              (bin, slice of <= 256 events) sums log(x), 1/x, (l/m)/x with x = 1 - (l/m) E/D
//              into per-bin fp64 accumulators.  The last CTA to finish computes G[b], F[b],
//              P[b] = sum_{b' <= b} G[b'], the loss and the header, and writes the (P,F) table.
//   K4 pass 2  (backward) streams the rows again, 9 B read + 4 B write:
//              grad = scale * (d - w * (P[b] - d * F[b]))
// Algorithmic HBM bytes: 22 per row for fwd+bwd; everything else is O(nbins).
#include <climits>

#include "common.cuh"

namespace b200surv {
namespace {

constexpr int P1_THREADS = 1024;
constexpr int P2_THREADS = 512;
constexpr int RED_THREADS = 1024;  // reduce: 32 bins x 32 groups of partials
constexpr int RED_BINS = 32;
constexpr int RED_NG = RED_THREADS / RED_BINS;
constexpr int RED_MAX_ITERS = 5;   // ceil(max pass-1 CTAs per segment / RED_NG): up to 160 CTAs
constexpr int MAX_P1_CTAS = RED_NG * RED_MAX_ITERS;
constexpr int IT_THREADS = 512;    // items / finish
constexpr int SLICE = 256;         // Efron events per warp task (8 per lane)
constexpr int MAX_PER = B200SURV_COX_MAX_BINS / IT_THREADS;  // bins per thread in the scans (16)
constexpr float LOG2E = 1.4426950408889634f;
constexpr double FIX_INV = 1.0 / 4294967296.0;   // 2^-32
constexpr float ETA_SCALE = 16777216.f;          // 2^24: fixed point of the sum of event log_hz
constexpr double ETA_INV = 1.0 / 16777216.0;
// shift is suitable when -16 <= max(log_hz) - shift <= 20 and the weights cannot overflow 2^31
constexpr float SHIFT_HI = 20.f, SHIFT_LO = -16.f;
constexpr double SUMW_LIMIT = 1073741824.0;      // 2^30

struct CtaRec {
    long long sum_ev_eta_q;  // sum over event rows of rint(log_hz * 2^24)
    double sum_w;            // estimate of the sum of weights (overflow guard)
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

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ================================================================ K1: pass 1
struct P1Acc {
    long long se_q;
    float mx, mt, sw;
    bool notbin, badt;
};

// smem layout: 5 words per bin, interleaved: [5*bin + {0: lo_cens, 1: hi_cens, 2: lo_event, 3: hi_event, 4: m}]
// (stride 5 is coprime with the 32 banks).  c2 = 32 - shift * log2(e): ex2(eta*log2e + c2) = w * 2^32.
__device__ __forceinline__ void p1_row(float eta, float t, bool ev, float c2, float c_sw, unsigned nb, unsigned *h,
                                       P1Acc &acc) {
    const float wq = ex2_approx(fmaf(eta, LOG2E, c2));
    const unsigned long long q = __float2ull_rn(wq);
    acc.mx = fmaxf(acc.mx, eta);
    acc.mt = fmaxf(acc.mt, t);
    acc.sw = fmaf(wq, c_sw, acc.sw);  // c_sw = 2^-32
    if (ev) acc.se_q += __float2ll_rn(eta * ETA_SCALE);
    int bin = __float2int_rz(t);
    const bool ok = ((unsigned)bin < nb) && ((float)bin == t);
    acc.notbin |= !ok;
    acc.badt |= !(t >= 0.f);
    bin = ok ? bin : 0;  // violating rows land in bin 0; the loss is poisoned through the flags anyway
    unsigned *base = h + 5 * bin + (ev ? 2 : 0);
    const unsigned lo = (unsigned)q;
    unsigned hi = (unsigned)(q >> 32);
    const unsigned old = atomicAdd(base, lo);
    hi += (old + lo < old);  // carry out of the low word
    if (hi) atomicAdd(base + 1, hi);
    if (ev) atomicAdd(h + 5 * bin + 4, 1u);
}

// partial layout per (seg, cta): u64 S_cens[nb], u64 S_event[nb], u32 m[nb]   (20 B/bin)
constexpr size_t PARTIAL_BYTES_PER_BIN = 20;

__global__ void __launch_bounds__(P1_THREADS, 1)
cox_binned_pass1(const float *__restrict__ log_hz, const float *__restrict__ time,
                 const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off, int64_t n,
                 int nb, float shift, int vec_ok, unsigned char *__restrict__ partial,
                 CtaRec *__restrict__ recs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *h = reinterpret_cast<unsigned *>(smem_raw);
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    __shared__ long long red_l[32];

    const int seg = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x;
    for (int i = threadIdx.x; i < 5 * nb; i += blockDim.x) h[i] = 0u;
    __syncthreads();

    const SegRange r = seg_range(seg_off, n, seg, vec_ok != 0);
    const float c2 = 32.f - shift * LOG2E;
    const float c_sw = 2.3283064365386963e-10f;
    const unsigned nbu = (unsigned)nb;
    P1Acc acc{0ll, -INFINITY, -INFINITY, 0.f, false, false};

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
        p1_row(e0.x, t0.x, (v0 & 0xffu) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.y, t0.y, (v0 & 0xff00u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.z, t0.z, (v0 & 0xff0000u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.w, t0.w, (v0 & 0xff000000u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e1.x, t1.x, (v1 & 0xffu) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e1.y, t1.y, (v1 & 0xff00u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e1.z, t1.z, (v1 & 0xff0000u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e1.w, t1.w, (v1 & 0xff000000u) != 0, c2, c_sw, nbu, h, acc);
    }
    if (g < ngroups) {
        const float4 e0 = ldg_stream_f4(lh + 4 * g);
        const float4 t0 = ldg_stream_f4(tm + 4 * g);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g);
        p1_row(e0.x, t0.x, (v0 & 0xffu) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.y, t0.y, (v0 & 0xff00u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.z, t0.z, (v0 & 0xff0000u) != 0, c2, c_sw, nbu, h, acc);
        p1_row(e0.w, t0.w, (v0 & 0xff000000u) != 0, c2, c_sw, nbu, h, acc);
    }
    // unaligned head [a, va) and tail [vb, b): scalar, spread over the CTAs
    {
        const int64_t nhead = r.va - r.a, ntail = r.b - r.vb;
        for (int64_t k = (int64_t)cta * blockDim.x + threadIdx.x; k < nhead + ntail; k += stride) {
            const int64_t row = (k < nhead) ? (r.a + k) : (r.vb + (k - nhead));
            p1_row(log_hz[row], time[row], event[row] != 0, c2, c_sw, nbu, h, acc);
        }
    }
    __syncthreads();

    // flush the CTA histogram with plain coalesced stores (summed exactly by the reduce kernel)
    unsigned char *out = partial + ((size_t)seg * nctas + cta) * PARTIAL_BYTES_PER_BIN * (size_t)nb;
    unsigned long long *o64 = reinterpret_cast<unsigned long long *>(out);
    unsigned *o32 = reinterpret_cast<unsigned *>(out + 16 * (size_t)nb);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const unsigned *hb = h + 5 * i;
        o64[i] = ((unsigned long long)hb[1] << 32) | hb[0];
        o64[nb + i] = ((unsigned long long)hb[3] << 32) | hb[2];
        o32[i] = hb[4];
    }
    const unsigned flags = (acc.notbin ? B200SURV_COXF_NOT_BINNABLE : 0u) | (acc.badt ? B200SURV_COXF_BAD_TIME : 0u);
    const double sw = block_reduce<double>((double)acc.sw, 0.0, OpAddD(), red_d);
    const long long se = block_reduce<long long>(acc.se_q, 0ll, OpAddLL(), red_l);
    const float mx = block_reduce<float>(acc.mx, -INFINITY, OpMaxF(), red_f);
    const float mt = block_reduce<float>(acc.mt, -INFINITY, OpMaxF(), red_f);
    const unsigned fl = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
    if (threadIdx.x == 0) {
        CtaRec rec;
        rec.sum_ev_eta_q = se; rec.sum_w = sw; rec.max_eta = mx; rec.max_time = mt; rec.flags = fl; rec.pad = 0;
        recs[(size_t)seg * nctas + cta] = rec;
    }
}

// ================================================================ K2: reduce partials
// bins layout per segment (int64): S_cens_q[nb], S_event_q[nb], m[nb], sum_ev_eta_q, n_not_binnable,
// ceil(sum_w), n_bad_time.  grid (nb / 32, n_seg), 1024 threads = 32 bins x 32 groups of partials.
// Also zeroes the Efron accumulators tgf[3][nb] and the ticket used by K3.
__global__ void __launch_bounds__(RED_THREADS)
cox_binned_reduce(const unsigned char *__restrict__ partial, const CtaRec *__restrict__ recs, int nctas, int nb,
                  long long *__restrict__ bins, float *__restrict__ bins_max, double *__restrict__ tgf_all,
                  unsigned *__restrict__ tickets) {
    __shared__ long long s_c[RED_NG][RED_BINS], s_e[RED_NG][RED_BINS];
    __shared__ unsigned s_m[RED_NG][RED_BINS];
    const int seg = blockIdx.y;
    const int lb = threadIdx.x & (RED_BINS - 1), grp = threadIdx.x / RED_BINS;
    const int b = blockIdx.x * RED_BINS + lb;
    long long *bs = bins + (size_t)seg * (3 * (size_t)nb + 4);
    {
        const unsigned char *p = partial + (size_t)seg * nctas * PARTIAL_BYTES_PER_BIN * (size_t)nb;
        unsigned long long vc[RED_MAX_ITERS], ve[RED_MAX_ITERS];
        unsigned vm[RED_MAX_ITERS];
#pragma unroll
        for (int k = 0; k < RED_MAX_ITERS; ++k) {  // all loads in flight together
            const int c = grp + k * RED_NG;
            const bool in = c < nctas;
            const unsigned char *pc = p + (size_t)(in ? c : 0) * PARTIAL_BYTES_PER_BIN * (size_t)nb;
            vc[k] = in ? reinterpret_cast<const unsigned long long *>(pc)[b] : 0ull;
            ve[k] = in ? reinterpret_cast<const unsigned long long *>(pc)[nb + b] : 0ull;
            vm[k] = in ? reinterpret_cast<const unsigned *>(pc + 16 * (size_t)nb)[b] : 0u;
        }
        unsigned long long sc = 0, se = 0;
        unsigned m = 0;
#pragma unroll
        for (int k = 0; k < RED_MAX_ITERS; ++k) { sc += vc[k]; se += ve[k]; m += vm[k]; }
        s_c[grp][lb] = (long long)sc; s_e[grp][lb] = (long long)se; s_m[grp][lb] = m;
    }
    if (threadIdx.x < 3 * RED_BINS) {
        double *tgf = tgf_all + (size_t)seg * 3 * nb;
        tgf[(threadIdx.x / RED_BINS) * nb + blockIdx.x * RED_BINS + lb] = 0.0;
    }
    __syncthreads();
    if (grp == 0) {
        long long sc = 0, se = 0, m = 0;
#pragma unroll
        for (int k = 0; k < RED_NG; ++k) { sc += s_c[k][lb]; se += s_e[k][lb]; m += s_m[k][lb]; }
        bs[b] = sc; bs[nb + b] = se; bs[2 * nb + b] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        long long se = 0;
        double sw = 0.0;
        float mx = -INFINITY, mt = -INFINITY;
        unsigned fl = 0;
        for (int c = threadIdx.x; c < nctas; c += 32) {
            const CtaRec r = recs[(size_t)seg * nctas + c];
            se += r.sum_ev_eta_q; sw += r.sum_w; mx = fmaxf(mx, r.max_eta); mt = fmaxf(mt, r.max_time); fl |= r.flags;
        }
        se = warp_sum(se); sw = warp_sum(sw); mx = warp_max(mx); mt = warp_max(mt); fl = warp_or(fl);
        if (threadIdx.x == 0) {
            bs[3 * (size_t)nb + 0] = se;
            bs[3 * (size_t)nb + 1] = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1 : 0;
            bs[3 * (size_t)nb + 2] = (long long)fmin(ceil(sw), 9.0e18);
            bs[3 * (size_t)nb + 3] = (fl & B200SURV_COXF_BAD_TIME) ? 1 : 0;
            bins_max[2 * seg + 0] = mx;
            bins_max[2 * seg + 1] = mt;
            tickets[seg] = 0;  // "last CTA done" ticket of the items kernel that follows
        }
    }
}

// ================================================================ block scans (512 threads)
// exclusive scan of one value per thread; returns the exclusive prefix, *total = block total
template <typename T>
__device__ __forceinline__ T block_exscan(T v, T *sh /*[33]*/, T *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = (lane < nw) ? sh[lane] : T(0), winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T u = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += u;
        }
        sh[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    *total = sh[32];
    return sh[wid] + (inc - v);
}

// ================================================================ K3: items + finish
// state per segment: header (64 B) | float2 (P, F)[nb]
__host__ __device__ inline size_t seg_state_stride(int nb) {
    return sizeof(b200surv_cox_header) + (size_t)nb * sizeof(float2);
}

// grid (gx, n_seg), 512 threads, dynamic smem: double D[nb] | int toff[nb + 1] | int m[nb]
template <int MAXPER>
__global__ void __launch_bounds__(IT_THREADS, MAXPER <= 8 ? 2 : 1)
cox_binned_items_finish(const long long *__restrict__ bins, const float *__restrict__ bins_max, int nb, int ties,
                        int reduction, float shift, double *tgf_all, unsigned *tickets,
                        float *__restrict__ out_loss, unsigned char *__restrict__ state) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sD = reinterpret_cast<double *>(smem_raw);
    int *s_toff = reinterpret_cast<int *>(smem_raw + sizeof(double) * nb);
    int *s_m = s_toff + (nb + 1);
    __shared__ double shd[33];
    __shared__ int shi[33];
    __shared__ int s_last;
    const int seg = blockIdx.y, t = threadIdx.x;
    const long long *bs = bins + (size_t)seg * (3 * (size_t)nb + 4);
    double *tgf = tgf_all + (size_t)seg * 3 * nb;
    const int per = nb / IT_THREADS > 0 ? nb / IT_THREADS : 1;  // nb >= 32: threads beyond nb idle
    const bool efron = ties == B200SURV_TIES_EFRON;

    // ---- redundant per-CTA scans of the nbins sums (all loads issued before any use)
    int n_events, n_tasks;
    {
        // suffix sums: thread t owns the reversed chunk [hi - per, hi)
        const int hi = nb - t * per;
        double sv[MAXPER];
        double loc = 0.0;
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = hi - 1 - k;
            const bool in = (k < per) && (b >= 0);
            sv[k] = in ? ((double)(unsigned long long)bs[b] + (double)(unsigned long long)bs[nb + b]) * FIX_INV : 0.0;
            loc += sv[k];
        }
        double tot;
        double run = block_exscan<double>(loc, shd, &tot);
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = hi - 1 - k;
            if (k < per && b >= 0) { run += sv[k]; sD[b] = run; }
        }
        // task offsets: exclusive prefix of ceil(m / SLICE), forward chunk [lo, lo + per)
        const int lo = t * per;
        int mv[MAXPER];
        int locm = 0, loct = 0;
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = lo + k;
            mv[k] = (k < per && b < nb) ? (int)bs[2 * nb + b] : 0;
            locm += mv[k];
            loct += (mv[k] + SLICE - 1) / SLICE;
        }
        int totm, tott;
        block_exscan<int>(locm, shi, &totm);
        int runt = block_exscan<int>(loct, shi, &tott);
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = lo + k;
            if (k < per && b < nb) { s_toff[b] = runt; s_m[b] = mv[k]; runt += (mv[k] + SLICE - 1) / SLICE; }
        }
        if (t == 0) s_toff[nb] = tott;
        n_events = totm; n_tasks = tott;
    }
    __syncthreads();

    // ---- Efron warp tasks: task k = slice (k - toff[b]) of bin b; 8 events per lane
    if (efron) {
        const int lane = t & 31, wpb = blockDim.x >> 5;
        for (int task = blockIdx.x * wpb + (t >> 5); task < n_tasks; task += gridDim.x * wpb) {
            int lo = 0, hi = nb - 1;  // first b with toff[b+1] > task
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_toff[mid + 1] > task) hi = mid; else lo = mid + 1;
            }
            const int b = lo;
            const int m = s_m[b];
            const double E = (double)(unsigned long long)bs[nb + b] * FIX_INV;
            const int l0 = (task - s_toff[b]) * SLICE, l1 = min(m, l0 + SLICE);
            const double rm = E / (sD[b] * (double)m);   // x_l = 1 - l * rm
            const float inv_m = 1.f / (float)m;
            float vt = 0.f, vg = 0.f, vf = 0.f;
            if ((double)(l1 - 1) * rm <= 0.5) {  // x >= 0.5: fp32 is accurate to ~1e-7 relative
                const float rmf = (float)rm;
#pragma unroll
                for (int j = 0; j < SLICE / 32; ++j) {
                    const int l = l0 + lane + 32 * j;
                    if (l < l1) {
                        const float x = fmaf(-(float)l, rmf, 1.f);
                        const float rx = rcp_approx(x);
                        vt += __logf(x);
                        vg += rx;
                        vf = fmaf((float)l * inv_m, rx, vf);
                    }
                }
            } else {  // the events are a large part of the risk set: keep the difference in fp64
                for (int l = l0 + lane; l < l1; l += 32) {
                    const double xd = 1.0 - (double)l * rm;
                    const float x = (float)xd;
                    const float rx = (float)(1.0 / xd);
                    vt += __logf(x);
                    vg += rx;
                    vf = fmaf((float)l * inv_m, rx, vf);
                }
            }
            vt = warp_sum(vt); vg = warp_sum(vg); vf = warp_sum(vf);
            if (lane == 0) {
                atomicAdd(tgf + b, (double)vt);
                atomicAdd(tgf + nb + b, (double)vg);
                atomicAdd(tgf + 2 * nb + b, (double)vf);
            }
        }
    }
    // ---- last CTA of the segment to arrive finishes
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(tickets + seg, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    unsigned char *seg_state = state + seg * seg_state_stride(nb);
    b200surv_cox_header *hdr = reinterpret_cast<b200surv_cox_header *>(seg_state);
    float2 *table = reinterpret_cast<float2 *>(seg_state + sizeof(b200surv_cox_header));
    {
        const int lo = t * per;
        double tv[MAXPER], gv[MAXPER], fv[MAXPER];
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {  // the accumulators were written by other CTAs: read through L2
            const int b = lo + k;
            const bool in = efron && (k < per) && (b < nb);
            tv[k] = in ? __ldcg(tgf + b) : 0.0;
            gv[k] = in ? __ldcg(tgf + nb + b) : 0.0;
            fv[k] = in ? __ldcg(tgf + 2 * nb + b) : 0.0;
        }
        double tsum = 0.0, gsum = 0.0;
        int net = 0;
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = lo + k;
            double G = 0.0, F = 0.0;
            if (k < per && b < nb) {
                const int m = s_m[b];
                if (m > 0) {
                    const double D = sD[b];
                    const double invD = 1.0 / D;
                    tsum += (double)m * log(D) + tv[k];
                    if (efron) { G = gv[k] * invD; F = fv[k] * invD; }
                    else G = (double)m * invD;
                    net += 1;
                }
            }
            gv[k] = G; fv[k] = F; gsum += G;
        }
        double tot;
        double run = block_exscan<double>(gsum, shd, &tot);
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = lo + k;
            if (k < per && b < nb) { run += gv[k]; table[b] = make_float2((float)run, (float)fv[k]); }
        }
        const double T = block_reduce<double>(tsum, 0.0, OpAddD(), shd);
        int n_times;
        block_exscan<int>(net, shi, &n_times);
        if (t == 0) {
            const double sum_ev_eta = (double)bs[3 * (size_t)nb] * ETA_INV;
            const double pll = sum_ev_eta - (T + (double)n_events * (double)shift);
            double norm = 1.0;
            if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)n_events;
            else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = efron ? (double)n_times : (double)n_events;
            unsigned flags = 0;
            if (bs[3 * (size_t)nb + 1] != 0) flags |= B200SURV_COXF_NOT_BINNABLE;
            if (bs[3 * (size_t)nb + 3] != 0) flags |= B200SURV_COXF_BAD_TIME;
            const float mx = bins_max[2 * seg];
            const double sumw = (double)bs[3 * (size_t)nb + 2];
            if (!(mx - shift <= SHIFT_HI) || !(mx - shift >= SHIFT_LO) || sumw >= SUMW_LIMIT)
                flags |= B200SURV_COXF_EXP_RANGE;
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
}

// ================================================================ K4: pass 2 (backward)
__device__ __forceinline__ float p2_row(float eta, float t, bool ev, float c2, float k, const float2 *tab, int nb) {
    const float w = ex2_approx(fmaf(eta, LOG2E, c2));  // c2 = -shift * log2(e)
    int bin = __float2int_rz(t);
    bin = min(max(bin, 0), nb - 1);  // invalid rows are already poisoned through k = NaN
    const float2 pf = tab[bin];
    const float d = ev ? 1.f : 0.f;
    return k * (d - w * (pf.x - d * pf.y));
}

__global__ void __launch_bounds__(P2_THREADS, 2)
cox_binned_bwd(const float *__restrict__ grad_out, const unsigned char *__restrict__ state,
               const float *__restrict__ log_hz, const float *__restrict__ time,
               const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off, int64_t n, int nb,
               int vec_ok, float *__restrict__ out_grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tab = reinterpret_cast<float2 *>(smem_raw);
    const int seg = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x, t = threadIdx.x;
    const unsigned char *seg_state = state + seg * seg_state_stride(nb);
    const b200surv_cox_header *hdr = reinterpret_cast<const b200surv_cox_header *>(seg_state);
    const float2 *gtab = reinterpret_cast<const float2 *>(seg_state + sizeof(b200surv_cox_header));
    for (int i = t; i < nb; i += blockDim.x) tab[i] = gtab[i];
    const float c2 = -hdr->shift * LOG2E;
    const float k = hdr->scale * grad_out[seg];
    __syncthreads();

    const SegRange r = seg_range(seg_off, n, seg, vec_ok != 0);
    const int64_t ngroups = (r.vb - r.va) >> 2;
    const int64_t stride = (int64_t)nctas * blockDim.x;
    const float *lh = log_hz + r.va;
    const float *tm = time + r.va;
    const uint8_t *evp = event + r.va;
    float *og = out_grad + r.va;
    int64_t g = (int64_t)cta * blockDim.x + t;
    for (; g + stride < ngroups; g += 2 * stride) {
        const int64_t g2 = g + stride;
        const float4 e0 = ldg_stream_f4(lh + 4 * g), e1 = ldg_stream_f4(lh + 4 * g2);
        const float4 t0 = ldg_stream_f4(tm + 4 * g), t1 = ldg_stream_f4(tm + 4 * g2);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g), v1 = ldg_stream_u32(evp + 4 * g2);
        float4 o0, o1;
        o0.x = p2_row(e0.x, t0.x, (v0 & 0xffu) != 0, c2, k, tab, nb);
        o0.y = p2_row(e0.y, t0.y, (v0 & 0xff00u) != 0, c2, k, tab, nb);
        o0.z = p2_row(e0.z, t0.z, (v0 & 0xff0000u) != 0, c2, k, tab, nb);
        o0.w = p2_row(e0.w, t0.w, (v0 & 0xff000000u) != 0, c2, k, tab, nb);
        o1.x = p2_row(e1.x, t1.x, (v1 & 0xffu) != 0, c2, k, tab, nb);
        o1.y = p2_row(e1.y, t1.y, (v1 & 0xff00u) != 0, c2, k, tab, nb);
        o1.z = p2_row(e1.z, t1.z, (v1 & 0xff0000u) != 0, c2, k, tab, nb);
        o1.w = p2_row(e1.w, t1.w, (v1 & 0xff000000u) != 0, c2, k, tab, nb);
        stg_stream_f4(og + 4 * g, o0);
        stg_stream_f4(og + 4 * g2, o1);
    }
    if (g < ngroups) {
        const float4 e0 = ldg_stream_f4(lh + 4 * g);
        const float4 t0 = ldg_stream_f4(tm + 4 * g);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g);
        float4 o0;
        o0.x = p2_row(e0.x, t0.x, (v0 & 0xffu) != 0, c2, k, tab, nb);
        o0.y = p2_row(e0.y, t0.y, (v0 & 0xff00u) != 0, c2, k, tab, nb);
        o0.z = p2_row(e0.z, t0.z, (v0 & 0xff0000u) != 0, c2, k, tab, nb);
        o0.w = p2_row(e0.w, t0.w, (v0 & 0xff000000u) != 0, c2, k, tab, nb);
        stg_stream_f4(og + 4 * g, o0);
    }
    {
        const int64_t nhead = r.va - r.a, ntail = r.b - r.vb;
        for (int64_t u = (int64_t)cta * blockDim.x + t; u < nhead + ntail; u += stride) {
            const int64_t row = (u < nhead) ? (r.a + u) : (r.vb + (u - nhead));
            out_grad[row] = p2_row(log_hz[row], time[row], event[row] != 0, c2, k, tab, nb);
        }
    }
}

// ================================================================ host-side layout
struct BinnedLayout {
    int nctas;  // pass-1 CTAs per segment
    size_t off_partial, off_recs, off_tickets, off_bins, off_bins_max, off_tgf, total;
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
    if (c > MAX_P1_CTAS) c = MAX_P1_CTAS;
    L.nctas = (int)c;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    L.off_partial = take((size_t)n_seg * L.nctas * PARTIAL_BYTES_PER_BIN * nb);
    L.off_recs = take((size_t)n_seg * L.nctas * sizeof(CtaRec));
    L.off_tickets = take((size_t)n_seg * sizeof(unsigned));
    L.off_bins = take((size_t)n_seg * (3 * (size_t)nb + 4) * sizeof(long long));
    L.off_bins_max = take((size_t)n_seg * 2 * sizeof(float));
    L.off_tgf = take((size_t)n_seg * 3 * nb * sizeof(double));
    L.total = o;
    return L;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int32_t check_common(int64_t n, int64_t n_seg, int nb) {
    B200_REQUIRE(nb >= 32 && nb <= B200SURV_COX_MAX_BINS && (nb & (nb - 1)) == 0,
                 "nbins must be a power of two in [32, 8192]");
    B200_REQUIRE(n >= 1 && n < (int64_t)1 << 31, "n must be in [1, 2^31)");
    B200_REQUIRE(n_seg >= 1 && n_seg <= 65535, "n_seg must be in [1, 65535]");
    return B200SURV_OK;
}

int32_t launch_pass1_reduce(const float *log_hz, const float *time, const uint8_t *event, const int64_t *seg_off,
                            int64_t n, int64_t n_seg, int nb, float shift, long long *bins, float *bins_max,
                            const BinnedLayout &L, unsigned char *w8, cudaStream_t st) {
    const int vec_ok = aligned16(log_hz) && aligned16(time) && ((reinterpret_cast<uintptr_t>(event) & 3) == 0);
    const size_t smem = (size_t)nb * PARTIAL_BYTES_PER_BIN;
    static bool attr_done = false;
    if (!attr_done) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(B200SURV_COX_MAX_BINS * PARTIAL_BYTES_PER_BIN)));
        attr_done = true;
    }
    unsigned char *partial = w8 + L.off_partial;
    CtaRec *recs = reinterpret_cast<CtaRec *>(w8 + L.off_recs);
    cox_binned_pass1<<<dim3(L.nctas, (unsigned)n_seg), P1_THREADS, smem, st>>>(log_hz, time, event, seg_off, n, nb,
                                                                               shift, vec_ok, partial, recs);
    cox_binned_reduce<<<dim3(nb / RED_BINS, (unsigned)n_seg), RED_THREADS, 0, st>>>(
        partial, recs, L.nctas, nb, bins, bins_max, reinterpret_cast<double *>(w8 + L.off_tgf),
        reinterpret_cast<unsigned *>(w8 + L.off_tickets));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t launch_items_finish(const long long *bins, const float *bins_max, int64_t n_seg, int ties, int reduction,
                            int nb, float shift, float *out_loss, void *state, const BinnedLayout &L,
                            unsigned char *w8, cudaStream_t st) {
    const size_t smem = (size_t)nb * sizeof(double) + (size_t)(2 * nb + 1) * sizeof(int);
    static bool attr_done = false;
    if (!attr_done) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_items_finish<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(B200SURV_COX_MAX_BINS * 16 + 16)));
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_items_finish<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(B200SURV_COX_MAX_BINS * 16 + 16)));
        attr_done = true;
    }
    int gx = 1;
    if (ties == B200SURV_TIES_EFRON) {
        gx = n_seg == 1 ? 2 * num_sms() : (int)((2 * num_sms() + n_seg - 1) / n_seg);
        if (gx < 1) gx = 1;
    }
    double *tgf = reinterpret_cast<double *>(w8 + L.off_tgf);
    unsigned *tickets = reinterpret_cast<unsigned *>(w8 + L.off_tickets);
    if (nb <= 8 * IT_THREADS)
        cox_binned_items_finish<8><<<dim3(gx, (unsigned)n_seg), IT_THREADS, smem, st>>>(
            bins, bins_max, nb, ties, reduction, shift, tgf, tickets, out_loss, static_cast<unsigned char *>(state));
    else
        cox_binned_items_finish<16><<<dim3(gx, (unsigned)n_seg), IT_THREADS, smem, st>>>(
            bins, bins_max, nb, ties, reduction, shift, tgf, tickets, out_loss, static_cast<unsigned char *>(state));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // namespace

// ================================================================ internal entry points
size_t cox_binned_state_bytes(int64_t n_seg, int nb) { return (size_t)n_seg * seg_state_stride(nb); }
size_t cox_binned_workspace_bytes(int64_t n, int64_t n_seg, int nb) { return binned_layout(n, n_seg, nb).total; }

// The workspace of cox_binned_finalize must be the one the preceding cox_binned_partial used (the reduce
// kernel zeroes the Efron accumulators and the ticket inside it).
int32_t cox_binned_partial(const float *log_hz, const float *time, const uint8_t *event,
                           const int64_t *seg_off, int64_t n, int64_t n_seg, int nb, float shift,
                           int64_t *bins_sum, float *bins_max, void *ws, size_t ws_bytes, cudaStream_t st) {
    int32_t rc = check_common(n, n_seg, nb);
    if (rc) return rc;
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    return launch_pass1_reduce(log_hz, time, event, seg_off, n, n_seg, nb, shift,
                               reinterpret_cast<long long *>(bins_sum), bins_max, L, static_cast<unsigned char *>(ws),
                               st);
}

int32_t cox_binned_finalize(const int64_t *bins_sum, const float *bins_max, int64_t n, int64_t n_seg, int ties,
                            int reduction, int nb, float shift, float *out_loss, void *state, size_t state_bytes,
                            void *ws, size_t ws_bytes, cudaStream_t st) {
    int32_t rc = check_common(n, n_seg, nb);
    if (rc) return rc;
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    if (state_bytes < cox_binned_state_bytes(n_seg, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    return launch_items_finish(reinterpret_cast<const long long *>(bins_sum), bins_max, n_seg, ties, reduction, nb,
                               shift, out_loss, state, L, static_cast<unsigned char *>(ws), st);
}

int32_t cox_binned_fwd(const float *log_hz, const float *time, const uint8_t *event, const int64_t *seg_off,
                       int64_t n, int64_t n_seg, int ties, int reduction, int nb, float shift,
                       float *out_loss, void *state, size_t state_bytes, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
    int32_t rc = check_common(n, n_seg, nb);
    if (rc) return rc;
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    if (state_bytes < cox_binned_state_bytes(n_seg, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    long long *bins = reinterpret_cast<long long *>(w8 + L.off_bins);
    float *bins_max = reinterpret_cast<float *>(w8 + L.off_bins_max);
    rc = launch_pass1_reduce(log_hz, time, event, seg_off, n, n_seg, nb, shift, bins, bins_max, L, w8, st);
    if (rc) return rc;
    return launch_items_finish(bins, bins_max, n_seg, ties, reduction, nb, shift, out_loss, state, L, w8, st);
}

int32_t cox_binned_bwd_launch(const float *grad_out, const void *state, size_t state_bytes, const float *log_hz,
                              const float *time, const uint8_t *event, const int64_t *seg_off, int64_t n,
                              int64_t n_seg, int nb, float *out_grad, cudaStream_t st) {
    int32_t rc = check_common(n, n_seg, nb);
    if (rc) return rc;
    if (state_bytes < cox_binned_state_bytes(n_seg, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
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
        grad_out, static_cast<const unsigned char *>(state), log_hz, time, event, seg_off, n, nb, vec_ok, out_grad);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // namespace b200surv
