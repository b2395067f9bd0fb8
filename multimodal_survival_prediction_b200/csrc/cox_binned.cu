// Cox negative partial log-likelihood, BINNED mode: integer-valued times in [0, nbins).
//
// Replaces torchsurv.loss.cox.neg_partial_log_likelihood as called by the reference at
// scripts/training/partial_modality_training.py:285-288 (math: oracle/cox.py header).
//
// Breslow/Efron depend on time only through per-distinct-time aggregates, so no sort is needed:
//   K1 pass 1  stream (log_hz,time,event) once, 9 B/row.  Per-CTA shared-memory histograms of
//              w = exp(log_hz - shift) in 36.28 FIXED POINT, accumulated with native 32-bit shared
//              atomics (low word; its carry and any integer part >= 16 go to the high word with a
//              second, rarely taken atomic) + an event counter per bin.  Integer accumulation is
//              exact and associative: the per-bin sums do not depend on the grid, the order of
//              the atomics or how rows are sharded over GPUs.
//              (fp32 atomicAdd in shared memory compiles to a CAS loop and is 7x slower --
//              measured, profiles/r1_hist_microbench.txt.)
//   K2 reduce  per-CTA partials -> per-bin int64 sums (this is what multi-GPU exchanges).
//   K3 finish  O(nbins): integer suffix sums D[b] = sum_{b' >= b} S[b'], the per-bin Breslow / Efron terms
//              (Efron's sums over the m tied events of a bin in closed form, Euler-Maclaurin: bin_terms),
//              P[b] = sum_{b' <= b} G[b'], the loss, the header and the (P,F) table.
//   K4 pass 2  (backward) streams the rows again, 9 B read + 4 B write:
//              grad = scale * (d - w * (P[b] - d * F[b]))
// One cohort per call (the headline case) runs K1..K3 as ONE cooperative kernel, cox_binned_fwd_fused: a single
// grid barrier after pass 1, then one warp per 32-bin block with decoupled look-back between the blocks; with
// peers it also carries the multi-GPU exchange of the per-bin sums over NVLink peer memory (PeerArgs).  Every
// path produces bit-identical results (integer sums; one canonical tree of floating-point additions).
// Algorithmic HBM bytes: 22 per row for fwd+bwd; everything else is O(nbins).
#include <cooperative_groups.h>

#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200surv {
namespace {

constexpr int P1_THREADS = 1024;
constexpr int P2_THREADS = 512;
constexpr int RED_THREADS = 1024;  // reduce: 32 bins x 32 groups of partials
constexpr int RED_BINS = 32;
constexpr int RED_NG = RED_THREADS / RED_BINS;
constexpr int RED_MAX_ITERS = 5;   // ceil(max pass-1 CTAs per segment / RED_NG): up to 160 CTAs
constexpr int MAX_P1_CTAS = RED_NG * RED_MAX_ITERS;
constexpr int IT_THREADS = 1024;   // finish
constexpr float LOG2E = 1.4426950408889634f;
constexpr int FIX_BITS = 28;
constexpr double FIX_INV = 1.0 / 268435456.0;    // 2^-28
constexpr double ETA_SCALE = 16777216.0;         // 2^24: fixed point of the sum of event log_hz
constexpr double ETA_INV = 1.0 / 16777216.0;
// shift is suitable when -8 <= max(log_hz) - shift <= 20 and the sum of weights stays below 2^30
constexpr float SHIFT_HI = 20.f, SHIFT_LO = -8.f;
constexpr double SUMW_LIMIT = 1073741824.0;      // 2^30
// a weight below 2^12 quanta (2^-16) carries a rounding error above 1e-4 of itself, and below half a quantum it is
// zero: min(log_hz) - shift < -16 ln 2 raises LOW_PRECISION (a late risk set made of such rows would be wrong)
constexpr float LOWP_MIN = -11.090354888959125f;

struct CtaRec {
    double sum_ev_eta;  // sum of log_hz over this CTA's event rows
    double sum_w;       // sum of weights (overflow guard)
    float max_eta, min_eta;
    unsigned flags, pad;
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
    float se, mx, mn, sw;  // sum of the event rows' log_hz, max / min log_hz, sum of the weights (overflow guard)
    bool notbin;   // a row was not an integer in [0, nbins) (also raised by NaN / negative times)
    bool badt;     // NaN or negative time: only looked for once notbin is up (cold path, see p1_bad_times)
};

// 64-bit adds into two 32-bit shared words with native atomics: low word (returning), its carry, then a predicated
// add of the high word; the event counter is a third predicated add (p1_prep / p1_atom_lo / p1_finish).
// smem layout: 5 words per bin, interleaved: [5*bin + {0: lo_cens, 1: hi_cens, 2: lo_event, 3: hi_event, 4: m}]
// (stride 5 is coprime with the 32 banks).  c2 = 28 - shift * log2(e): ex2(eta*log2e + c2) = w * 2^28.
struct P1Row {
    uint32_t a_bin;  // shared address of the bin's 5 words
    unsigned lo, hi;
    bool ev;
};
__device__ __forceinline__ P1Row p1_prep(float eta, float t, bool ev, float c2, unsigned nb, uint32_t h_addr,
                                         P1Acc &acc) {
    const float wq = ex2_approx(fmaf(eta, LOG2E, c2));
    const unsigned long long q = __float2ull_rn(wq);
    acc.mx = fmaxf(acc.mx, eta);
    acc.mn = fminf(acc.mn, eta);
    acc.sw += wq;
    acc.se += ev ? eta : 0.f;
    int bin = __float2int_rz(t);
    const bool ok = ((unsigned)bin < nb) && ((float)bin == t);
    acc.notbin |= !ok;
    bin = ok ? bin : 0;  // violating rows land in bin 0; the loss is poisoned through the flags anyway
    P1Row r;
    r.a_bin = h_addr + 20u * (unsigned)bin;
    r.lo = (unsigned)q; r.hi = (unsigned)(q >> 32); r.ev = ev;
    return r;
}
__device__ __forceinline__ unsigned p1_atom_lo(const P1Row &r) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(r.a_bin + (r.ev ? 8u : 0u)), "r"(r.lo) : "memory");
    return old;
}
__device__ __forceinline__ void p1_finish(const P1Row &r, unsigned old) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, pe;\n\t"
        ".reg .u32 s, h;\n\t"
        "add.cc.u32 s, %1, %2;\n\t"
        "addc.u32 h, %3, 0;\n\t"
        "setp.ne.u32 p, h, 0;\n\t"
        "@p red.shared.add.u32 [%0+4], h;\n\t"
        "setp.ne.u32 pe, %4, 0;\n\t"
        "@pe red.shared.add.u32 [%5], 1;\n\t"
        "}\n" ::"r"(r.a_bin + (r.ev ? 8u : 0u)), "r"(old), "r"(r.lo), "r"(r.hi), "r"((unsigned)r.ev), "r"(r.a_bin + 16u)
        : "memory");
}
// BAD_TIME is only reported next to NOT_BINNABLE (every NaN / negative time is not binnable): out of line, so that
// the hot loop pays one never-taken branch per group instead of a compare per row
__device__ __noinline__ bool p1_bad_times(float a, float b, float c, float d) {
    return !(a >= 0.f) || !(b >= 0.f) || !(c >= 0.f) || !(d >= 0.f);
}
// four rows of one 128-bit group: the four returning atomics are issued back to back, then the carries
__device__ __forceinline__ void p1_group(const float4 e, const float4 t, uint32_t v, float c2, unsigned nb,
                                         uint32_t h_addr, P1Acc &acc) {
    const P1Row r0 = p1_prep(e.x, t.x, (v & 0xffu) != 0, c2, nb, h_addr, acc);
    const P1Row r1 = p1_prep(e.y, t.y, (v & 0xff00u) != 0, c2, nb, h_addr, acc);
    const P1Row r2 = p1_prep(e.z, t.z, (v & 0xff0000u) != 0, c2, nb, h_addr, acc);
    const P1Row r3 = p1_prep(e.w, t.w, (v & 0xff000000u) != 0, c2, nb, h_addr, acc);
    const unsigned o0 = p1_atom_lo(r0), o1 = p1_atom_lo(r1), o2 = p1_atom_lo(r2), o3 = p1_atom_lo(r3);
    p1_finish(r0, o0); p1_finish(r1, o1); p1_finish(r2, o2); p1_finish(r3, o3);
    if (acc.notbin) acc.badt |= p1_bad_times(t.x, t.y, t.z, t.w);
}
__device__ __forceinline__ void p1_row(float eta, float t, bool ev, float c2, unsigned nb, uint32_t h_addr,
                                       P1Acc &acc) {
    const P1Row r = p1_prep(eta, t, ev, c2, nb, h_addr, acc);
    p1_finish(r, p1_atom_lo(r));
    if (acc.notbin) acc.badt |= p1_bad_times(t, 0.f, 0.f, 0.f);
}

// partial layout per (seg, cta): u64 S_cens[nb], u64 S_event[nb], u32 m[nb]   (20 B/bin)
constexpr size_t PARTIAL_BYTES_PER_BIN = 20;

// end of pass 1: flush the CTA histogram with plain coalesced stores (summed exactly by the reduce step) and
// write the CTA record.  All threads of the block must call.
__device__ __forceinline__ void pass1_flush(const unsigned *h, const P1Acc &acc, double se_d, double sw_d,
                                            unsigned char *__restrict__ partial, CtaRec *__restrict__ recs, int seg,
                                            int cta, int nctas, int nb, double *red_d, float *red_f, unsigned *red_u) {
    __syncthreads();
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
    const double sw = block_reduce<double>(sw_d * FIX_INV, 0.0, OpAddD(), red_d);
    const double se = block_reduce<double>(se_d, 0.0, OpAddD(), red_d);
    const float mx = block_reduce<float>(acc.mx, -INFINITY, OpMaxF(), red_f);
    const float mn = -block_reduce<float>(-acc.mn, -INFINITY, OpMaxF(), red_f);
    const unsigned fl = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
    if (threadIdx.x == 0) {
        CtaRec rec;
        rec.sum_ev_eta = se; rec.sum_w = sw; rec.max_eta = mx; rec.min_eta = mn; rec.flags = fl; rec.pad = 0;
        recs[(size_t)seg * nctas + cta] = rec;
    }
}

// body of pass 1 for CTA `cta` of `nctas` of segment `seg` (shared by the stand-alone and the fused kernel)
__device__ __forceinline__ void pass1_body(const float *__restrict__ log_hz, const float *__restrict__ time,
                                           const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off,
                                           int64_t n, int nb, float shift, int vec_ok,
                                           unsigned char *__restrict__ partial, CtaRec *__restrict__ recs,
                                           int seg, int cta, int nctas, unsigned *h) {
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    for (int i = threadIdx.x; i < 5 * nb; i += blockDim.x) h[i] = 0u;
    __syncthreads();

    const SegRange r = seg_range(seg_off, n, seg, vec_ok != 0);
    const float c2 = (float)FIX_BITS - shift * LOG2E;
    const unsigned nbu = (unsigned)nb;
    const uint32_t h_addr = (uint32_t)__cvta_generic_to_shared(h);
    P1Acc acc{0.f, -INFINITY, INFINITY, 0.f, false, false};
    double se_d = 0.0, sw_d = 0.0;  // per-thread fp32 partials are folded into fp64 every iteration

    // 128-bit groups, two per thread per iteration (all six loads issued before any use)
    const int64_t ngroups = (r.vb - r.va) >> 2;
    const int64_t stride = (int64_t)nctas * blockDim.x;
    int64_t g = (int64_t)cta * blockDim.x + threadIdx.x;
    const float *lh = log_hz + r.va;
    const float *tm = time + r.va;
    const uint8_t *evp = event + r.va;
    // (a register-rolling prefetch of the next iteration was measured slower here: the 64-register
    // cap of a 1024-thread CTA makes it spill -- profiles/r1_v7_launches.csv)
    for (; g + stride < ngroups; g += 2 * stride) {
        const int64_t g2 = g + stride;
        const float4 e0 = ldg_stream_f4(lh + 4 * g), e1 = ldg_stream_f4(lh + 4 * g2);
        const float4 t0 = ldg_stream_f4(tm + 4 * g), t1 = ldg_stream_f4(tm + 4 * g2);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g), v1 = ldg_stream_u32(evp + 4 * g2);
        p1_group(e0, t0, v0, c2, nbu, h_addr, acc);
        p1_group(e1, t1, v1, c2, nbu, h_addr, acc);
        se_d += (double)acc.se; sw_d += (double)acc.sw; acc.se = 0.f; acc.sw = 0.f;
    }
    if (g < ngroups) {
        const float4 e0 = ldg_stream_f4(lh + 4 * g);
        const float4 t0 = ldg_stream_f4(tm + 4 * g);
        const uint32_t v0 = ldg_stream_u32(evp + 4 * g);
        p1_group(e0, t0, v0, c2, nbu, h_addr, acc);
    }
    // unaligned head [a, va) and tail [vb, b): scalar, spread over the CTAs
    {
        const int64_t nhead = r.va - r.a, ntail = r.b - r.vb;
        for (int64_t k = (int64_t)cta * blockDim.x + threadIdx.x; k < nhead + ntail; k += stride) {
            const int64_t row = (k < nhead) ? (r.a + k) : (r.vb + (k - nhead));
            p1_row(log_hz[row], time[row], event[row] != 0, c2, nbu, h_addr, acc);
        }
    }
    se_d += (double)acc.se; sw_d += (double)acc.sw;
    pass1_flush(h, acc, se_d, sw_d, partial, recs, seg, cta, nctas, nb, red_d, red_f, red_u);
}

// TMA-staged variant of the pass-1 body (one cohort, 16-byte aligned inputs, nbins <= 4096): warp 31 is a
// producer that keeps TMA_STAGES bulk copies (cp.async.bulk, mbarrier tx-count) of 3968-row tiles in flight;
// the other 31 warps bin the rows out of shared memory.  Memory latency is hidden by the ring instead of by
// registers (a 1024-thread CTA only has 64 registers per thread).
constexpr int TMA_STAGES = 3;
constexpr int TMA_CONSUMERS = P1_THREADS - 32;
constexpr int TMA_TILE = TMA_CONSUMERS * 4;       // 3968 rows
constexpr int TMA_STAGE_BYTES = TMA_TILE * 9;     // eta f32 | time f32 | event u8 = 35,712 B
__host__ __device__ inline size_t tma_hist_bytes(int nb) { return ((size_t)nb * 24 + 16 + 127) / 128 * 128; }
__host__ __device__ inline size_t tma_smem_bytes(int nb) {
    return tma_hist_bytes(nb) + (size_t)TMA_STAGES * TMA_STAGE_BYTES + 2 * TMA_STAGES * sizeof(uint64_t);
}

__device__ __forceinline__ void pass1_body_tma(const float *__restrict__ log_hz, const float *__restrict__ time,
                                               const uint8_t *__restrict__ event, int64_t n, int nb, float shift,
                                               unsigned char *__restrict__ partial, CtaRec *__restrict__ recs,
                                               int cta, int nctas, unsigned char *smem_raw) {
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    unsigned *h = reinterpret_cast<unsigned *>(smem_raw);
    unsigned char *stages = smem_raw + tma_hist_bytes(nb);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + (size_t)TMA_STAGES * TMA_STAGE_BYTES);  // full[S], empty[S]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 5 * nb; i += blockDim.x) h[i] = 0u;
    if (t == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbarrier_init(smem_addr_u32(bars + s), 1);
            mbarrier_init(smem_addr_u32(bars + TMA_STAGES + s), TMA_CONSUMERS / 32);
        }
        mbarrier_init_fence();
    }
    __syncthreads();

    const float c2 = (float)FIX_BITS - shift * LOG2E;
    const unsigned nbu = (unsigned)nb;
    const uint32_t h_addr = smem_addr_u32(h);
    P1Acc acc{0.f, -INFINITY, INFINITY, 0.f, false, false};
    double se_d = 0.0, sw_d = 0.0;
    const int64_t ntiles = n / TMA_TILE;  // full tiles; the remainder goes through direct loads below
    if (warp == P1_THREADS / 32 - 1) {
        if (lane == 0) {  // producer
            int s = 0;
            uint32_t ph = 0;
            for (int64_t tile = cta; tile < ntiles; tile += nctas) {
                mbarrier_wait(smem_addr_u32(bars + TMA_STAGES + s), ph ^ 1);
                const uint32_t full = smem_addr_u32(bars + s);
                mbarrier_expect_tx(full, TMA_STAGE_BYTES);
                const uint32_t dst = smem_addr_u32(stages + (size_t)s * TMA_STAGE_BYTES);
                const int64_t row0 = tile * TMA_TILE;
                bulk_load_1d(dst, log_hz + row0, TMA_TILE * 4, full);
                bulk_load_1d(dst + TMA_TILE * 4, time + row0, TMA_TILE * 4, full);
                bulk_load_1d(dst + TMA_TILE * 8, event + row0, TMA_TILE, full);
                if (++s == TMA_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {  // consumers: thread t bins rows 4t .. 4t+3 of every tile
        int s = 0;
        uint32_t ph = 0;
        for (int64_t tile = cta; tile < ntiles; tile += nctas) {
            mbarrier_wait(smem_addr_u32(bars + s), ph);
            const unsigned char *st = stages + (size_t)s * TMA_STAGE_BYTES;
            const float4 e4 = *reinterpret_cast<const float4 *>(st + 16 * t);
            const float4 t4 = *reinterpret_cast<const float4 *>(st + TMA_TILE * 4 + 16 * t);
            const uint32_t v4 = *reinterpret_cast<const uint32_t *>(st + TMA_TILE * 8 + 4 * t);
            __syncwarp();
            if (lane == 0) mbarrier_arrive(smem_addr_u32(bars + TMA_STAGES + s));  // data is in registers: free the slot
            p1_group(e4, t4, v4, c2, nbu, h_addr, acc);
            se_d += (double)acc.se; sw_d += (double)acc.sw; acc.se = 0.f; acc.sw = 0.f;
            if (++s == TMA_STAGES) { s = 0; ph ^= 1; }
        }
    }
    {  // remainder rows (fewer than one tile): direct loads, spread over the whole grid
        const int64_t rem0 = ntiles * TMA_TILE;
        for (int64_t row = rem0 + (int64_t)cta * blockDim.x + t; row < n; row += (int64_t)nctas * blockDim.x)
            p1_row(log_hz[row], time[row], event[row] != 0, c2, nbu, h_addr, acc);
    }
    se_d += (double)acc.se; sw_d += (double)acc.sw;
    pass1_flush(h, acc, se_d, sw_d, partial, recs, 0, cta, nctas, nb, red_d, red_f, red_u);
}

// cp.async ring variant of the pass-1 body (one cohort, 16-byte aligned inputs, nbins <= 4096): every thread
// keeps RING_STAGES - 1 iterations of its OWN three loads (16 B log_hz, 16 B time, 4 B event) in flight as
// asynchronous global->shared copies and bins the oldest one.  A thread only ever reads back what it copied
// itself, so cp.async.wait_group is the only synchronisation (no barrier, no producer warp).
constexpr int RING_STAGES = 3;
constexpr int RING_STAGE_BYTES = P1_THREADS * 36;  // 36,864 B
__host__ __device__ inline size_t ring_smem_bytes(int nb) { return tma_hist_bytes(nb) + (size_t)RING_STAGES * RING_STAGE_BYTES; }

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void pass1_body_ring(const float *__restrict__ log_hz, const float *__restrict__ time,
                                                const uint8_t *__restrict__ event, int64_t n, int nb, float shift,
                                                unsigned char *__restrict__ partial, CtaRec *__restrict__ recs,
                                                int cta, int nctas, unsigned char *smem_raw) {
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    unsigned *h = reinterpret_cast<unsigned *>(smem_raw);
    const int t = threadIdx.x;
    for (int i = t; i < 5 * nb; i += blockDim.x) h[i] = 0u;
    __syncthreads();

    const float c2 = (float)FIX_BITS - shift * LOG2E;
    const unsigned nbu = (unsigned)nb;
    const uint32_t h_addr = smem_addr_u32(h);
    const uint32_t ring = smem_addr_u32(smem_raw + tma_hist_bytes(nb));
    const uint32_t my_e = ring + 16u * t, my_t = ring + 16u * P1_THREADS + 16u * t, my_v = ring + 32u * P1_THREADS + 4u * t;
    P1Acc acc{0.f, -INFINITY, INFINITY, 0.f, false, false};
    double se_d = 0.0, sw_d = 0.0;
    const int64_t ngroups = n >> 2, stride = (int64_t)nctas * P1_THREADS;
    const int64_t g0 = (int64_t)cta * P1_THREADS + t;
    const int iters = g0 < ngroups ? (int)((ngroups - g0 + stride - 1) / stride) : 0;
    // running source pointers and a countdown instead of 64-bit index arithmetic per copy (the loop is bound by
    // instruction issue: ~45 instructions per row at 2.8 IPC)
    const char *pe = reinterpret_cast<const char *>(log_hz + 4 * g0), *pt = reinterpret_cast<const char *>(time + 4 * g0);
    const char *pv = reinterpret_cast<const char *>(event + 4 * g0);
    const int64_t step16 = 16 * stride, step4 = 4 * stride;
    int to_issue = iters;
    uint32_t so_issue = 0;
    auto issue = [&]() {
        if (to_issue > 0) {
            cp_async_16(my_e + so_issue, pe);
            cp_async_16(my_t + so_issue, pt);
            cp_async_4(my_v + so_issue, pv);
            pe += step16; pt += step16; pv += step4;
            --to_issue;
        }
        cp_async_commit();
        so_issue += RING_STAGE_BYTES;
        if (so_issue == RING_STAGES * RING_STAGE_BYTES) so_issue = 0;
    };
#pragma unroll
    for (int s = 0; s < RING_STAGES - 1; ++s) issue();
    uint32_t so = 0;
    for (int i = 0; i < iters; ++i) {
        issue();
        cp_async_wait<RING_STAGES - 1>();
        float4 e4, t4;
        uint32_t v4;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e4.x), "=f"(e4.y), "=f"(e4.z), "=f"(e4.w) : "r"(my_e + so));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w) : "r"(my_t + so));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v4) : "r"(my_v + so));
        p1_group(e4, t4, v4, c2, nbu, h_addr, acc);
        se_d += (double)acc.se; sw_d += (double)acc.sw; acc.se = 0.f; acc.sw = 0.f;
        so += RING_STAGE_BYTES;
        if (so == RING_STAGES * RING_STAGE_BYTES) so = 0;
    }
    cp_async_wait<0>();
    {  // rows beyond the last full 4-row group: direct loads
        for (int64_t row = (ngroups << 2) + (int64_t)cta * blockDim.x + t; row < n; row += stride)
            p1_row(log_hz[row], time[row], event[row] != 0, c2, nbu, h_addr, acc);
    }
    se_d += (double)acc.se; sw_d += (double)acc.sw;
    pass1_flush(h, acc, se_d, sw_d, partial, recs, 0, cta, nctas, nb, red_d, red_f, red_u);
}

__global__ void __launch_bounds__(P1_THREADS, 1)
cox_binned_pass1(const float *__restrict__ log_hz, const float *__restrict__ time,
                 const uint8_t *__restrict__ event, const int64_t *__restrict__ seg_off, int64_t n,
                 int nb, float shift, int vec_ok, unsigned char *__restrict__ partial,
                 CtaRec *__restrict__ recs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pass1_body(log_hz, time, event, seg_off, n, nb, shift, vec_ok, partial, recs, blockIdx.y, blockIdx.x, gridDim.x,
               reinterpret_cast<unsigned *>(smem_raw));
}

// ================================================================ per-bin terms (Breslow / Efron)
// A bin with m tied events, event weight E and risk-set weight D contributes
//   T = m log D + sum_{l<m} log x_l,   G = (1/D) sum_{l<m} 1/x_l,   F = (1/D) sum_{l<m} (l/m)/x_l,   x_l = 1 - (l/m) E/D
// (Breslow: x_l = 1).  The three sums over l are smooth functions of l sampled at the integers, so they are
// evaluated in O(1) per bin with the Euler-Maclaurin formula (integral + end-point term + five Bernoulli
// corrections) over l in [0, L), where L is the largest index that keeps the derivative ratio
// a / (1 - a L) <= 1/EM_Q, a = E/(D m): the first omitted correction is then < 1e-12 of the sum.  The at most
// EM_Q + 1 remaining terms (risk sets that are almost only the tied events) are summed directly.  The whole tail of
// the forward pass is therefore O(nbins), independent of the number of events (scratch/em_efron.py checks the
// expansion against direct long-double summation: relative error <= 1e-11 for m up to 3e6 and any E/D in [0,1]).
constexpr int EM_Q = 16;
struct BinTerms { double T, G, F; };

// (one out-of-line copy of log keeps the tail's code small)
__device__ __noinline__ double log_f64(double x) { return log(x); }

__device__ __noinline__ BinTerms bin_terms(long long Dq, long long Eq, int m, int efron) {
    BinTerms o;
    o.T = 0.0; o.G = 0.0; o.F = 0.0;
    const bool empty = m <= 0;  // (no early return: the lanes of a block stay together; an empty bin computes on 1, 1)
    if (empty) { m = 1; Dq = 1ll << FIX_BITS; Eq = 0; }
    const double md = (double)m, invm = 1.0 / md;
    const double D = (double)Dq * FIX_INV, invD = 1.0 / D;
    double T = md * log_f64(D);
    if (!efron) { o.T = empty ? 0.0 : T; o.G = empty ? 0.0 : md * invD; return o; }
    const double r = ((double)Eq * FIX_INV) * invD;  // in [0, 1]: the bin's events are part of its risk set
    const double a = r * invm;
    int L = m;
    if (1.0 - r < (double)EM_Q * a) {
        const double Lf = floor(1.0 / a - (double)EM_Q);
        L = Lf < 2.0 ? 0 : (int)fmin(Lf, md);
    }
    if (L < 2) L = 0;
    double sg = 0.0, sf = 0.0;
    if (L > 0) {
        const double Ld = (double)L, z = a * Ld;  // z <= 1 - EM_Q a < 1
        // ps = -log(1-z)/z, ph = (ps - 1)/z, lz = log(1-z); power series where the closed forms cancel
        double ps, ph, lz;
        if (z < 0.03125) {
            ps = 1.0 / 14.0; ph = 1.0 / 15.0;
#pragma unroll
            for (int j = 12; j >= 0; --j) { ps = fma(ps, z, 1.0 / (double)(j + 1)); ph = fma(ph, z, 1.0 / (double)(j + 2)); }
            lz = -z * ps;
        } else {
            const double invz = 1.0 / z;
            lz = log_f64(1.0 - z);
            ps = -lz * invz;
            ph = (ps - 1.0) * invz;
        }
        const double u = 1.0 / (1.0 - z), u2 = u * u, a2 = a * a;
        double st = -Ld * z * (ps - ph) - 0.5 * lz;
        sg = Ld * ps + 0.5 * (1.0 - u);
        sf = (Ld * Ld * invm) * ph - 0.5 * (Ld * invm) * u;
        // Bernoulli corrections k = 1..5:  -B2k/(2k(2k-1)) for the logs,  B2k/(2k) for the reciprocals
        const double CL[5] = {-1.0 / 12.0, 1.0 / 360.0, -1.0 / 1260.0, 1.0 / 1680.0, -1.0 / 1188.0};
        const double CI[5] = {1.0 / 12.0, -1.0 / 120.0, 1.0 / 252.0, -1.0 / 240.0, 1.0 / 132.0};
        double pw_odd = u, pw_even = u2, ak = a, akm = invm;  // (1-z)^-(2k-1), (1-z)^-2k, a^(2k-1), a^(2k-2)/m
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            st = fma(CL[k] * ak, pw_odd - 1.0, st);
            sg = fma(CI[k] * ak, pw_even - 1.0, sg);
            sf = fma(CI[k] * akm, pw_even - 1.0, sf);
            pw_odd *= u2; pw_even *= u2; ak *= a2; akm *= a2;
        }
        T += st;
    }
    if (L < m) {  // at most EM_Q + 1 terms; the logs as the log of the product (every factor >= 1/m)
        double prod = 1.0;
#pragma unroll 1
        for (int l = L; l < m; ++l) {
            const double x = 1.0 - (double)l * a, rx = 1.0 / x;
            prod *= x;
            sg += rx;
            sf = fma((double)l * invm, rx, sf);
        }
        T += log_f64(prod);
    }
    o.T = empty ? 0.0 : T; o.G = empty ? 0.0 : sg * invD; o.F = empty ? 0.0 : sf * invD;
    return o;
}

// ================================================================ canonical warp scans
// P[b] = sum_{b' <= b} G[b'] and the total of T are formed by the SAME tree of additions on every code path
// (fused forward on any grid, multi-GPU, stand-alone finish kernel), so that the loss and the (P,F) table -- and
// with them every gradient -- are bit-identical across the paths: level 1 is a Hillis-Steele scan over the 32
// bins of a block, level 2 the same scan over the 32 block totals of a superblock (1024 bins), level 3 a
// sequential sum over the superblocks.  The suffix sums D are integers (any order is exact).
__device__ __forceinline__ double hs_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += u;
    }
    return v;
}
__device__ __forceinline__ long long suffix_scan(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_down_sync(FULL, v, o);
        if (lane + o < 32) v += u;
    }
    return v;
}

// ================================================================ K2: reduce partials
// grid (nb / 32, n_seg), 1024 threads = 32 bins x 32 groups of partials.
__global__ void __launch_bounds__(RED_THREADS)
cox_binned_reduce(const unsigned char *__restrict__ partial, const CtaRec *__restrict__ recs, int nctas, int nb,
                  long long *bins, float *__restrict__ bins_max) {
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
    __syncthreads();
    if (grp == 0) {
        long long sc = 0, se = 0, m = 0;
#pragma unroll
        for (int k = 0; k < RED_NG; ++k) { sc += s_c[k][lb]; se += s_e[k][lb]; m += s_m[k][lb]; }
        bs[b] = sc; bs[nb + b] = se; bs[2 * nb + b] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double se = 0.0, sw = 0.0;
        float mx = -INFINITY, nmn = -INFINITY;
        unsigned fl = 0;
        for (int c = threadIdx.x; c < nctas; c += 32) {
            const CtaRec r = recs[(size_t)seg * nctas + c];
            se += r.sum_ev_eta; sw += r.sum_w; mx = fmaxf(mx, r.max_eta); nmn = fmaxf(nmn, -r.min_eta); fl |= r.flags;
        }
        se = warp_sum(se); sw = warp_sum(sw); mx = warp_max(mx); nmn = warp_max(nmn); fl = warp_or(fl);
        if (threadIdx.x == 0) {
            bs[3 * (size_t)nb + 0] = __double2ll_rn(se * ETA_SCALE);
            bs[3 * (size_t)nb + 1] = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1 : 0;
            bs[3 * (size_t)nb + 2] = (long long)fmin(ceil(sw), 9.0e18);
            bs[3 * (size_t)nb + 3] = (fl & B200SURV_COXF_BAD_TIME) ? 1 : 0;
            bins_max[2 * seg + 0] = mx;
            bins_max[2 * seg + 1] = nmn;  // MINUS the smallest log_hz (so that a MAX all-reduce combines both words)
        }
    }
}

// ================================================================ header
// state per segment: header (64 B) | float2 (P, F)[nb]
__host__ __device__ inline size_t seg_state_stride(int nb) {
    return sizeof(b200surv_cox_header) + (size_t)nb * sizeof(float2);
}

struct HeaderIn {
    long long sum_ev_eta_q, n_not_binnable, sum_w_ceil, n_bad_time;  // the four scalar words of the per-bin sums
    float max_eta, min_eta;
    long long n_events, n_times;
    double T;  // sum over the bins of m log D + sum_l log x_l
    int peer_timeout;
};
__device__ __forceinline__ void write_header(const HeaderIn &h, int efron, int reduction, float shift, int nb,
                                             b200surv_cox_header *hdr, float *out_loss) {
    const double sum_ev_eta = (double)h.sum_ev_eta_q * ETA_INV;
    const double pll = sum_ev_eta - (h.T + (double)h.n_events * (double)shift);
    double norm = 1.0;
    if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)h.n_events;
    else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = efron ? (double)h.n_times : (double)h.n_events;
    unsigned flags = 0;
    if (h.n_not_binnable != 0) flags |= B200SURV_COXF_NOT_BINNABLE;
    if (h.n_bad_time != 0) flags |= B200SURV_COXF_BAD_TIME;
    const float mx = h.max_eta;
    if (!(mx - shift <= SHIFT_HI) || !(mx - shift >= SHIFT_LO) || (double)h.sum_w_ceil >= SUMW_LIMIT)
        flags |= B200SURV_COXF_EXP_RANGE;
    if (h.peer_timeout) flags |= B200SURV_COXF_PEER_TIMEOUT;
    if (h.min_eta - shift < LOWP_MIN) flags |= B200SURV_COXF_LOW_PRECISION;
    float loss = 0.f, scale = 0.f;
    if (h.n_events > 0) { loss = (float)(-pll / norm); scale = (float)(-1.0 / norm); }
    if (flags) { loss = __int_as_float(0x7fc00000); scale = loss; }
    hdr->flags = flags; hdr->mode = B200SURV_COX_BINNED; hdr->loss = loss; hdr->scale = scale;
    hdr->shift = shift; hdr->max_log_hz = mx; hdr->max_time = -1.f;
    hdr->nbins = nb; hdr->n_events = h.n_events; hdr->n_event_times = h.n_times; hdr->pll = pll;
    hdr->min_log_hz = h.min_eta; hdr->reserved = 0;
    *out_loss = loss;
}

// ================================================================ K3: finish (stand-alone: segmented cohorts, NCCL path)
// One CTA of 1024 threads per cohort, straight from the per-bin int64 sums: suffix sums D, per-bin terms, prefix
// sums P, loss, header and the (P,F) table.  Warp w owns the 32-bin blocks w, w + 32, ...
__global__ void __launch_bounds__(IT_THREADS, 1)
cox_binned_finish(const long long *__restrict__ bins, const float *__restrict__ bins_max, int nb, int ties,
                  int reduction, float shift, double *__restrict__ scr_g, double *__restrict__ scr_f,
                  float *__restrict__ out_loss, unsigned char *__restrict__ state) {
    constexpr int MAXBLK = B200SURV_COX_MAX_BINS / 32;
    __shared__ long long s_tot[MAXBLK], s_off[MAXBLK];
    __shared__ double s_g[MAXBLK], s_t[MAXBLK];
    __shared__ int s_ne, s_nt;
    const int seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long *bs = bins + (size_t)seg * (3 * (size_t)nb + 4);
    double *sg = scr_g + (size_t)seg * nb, *sf = scr_f + (size_t)seg * nb;
    const bool efron = ties == B200SURV_TIES_EFRON;
    const int nblk = nb >> 5, nsup = (nblk + 31) >> 5;
    if (t == 0) { s_ne = 0; s_nt = 0; }
    __syncthreads();
    for (int j = 0; j < nsup; ++j) {  // block totals of S = S_cens + S_event, event counts
        const int k = warp + 32 * j;
        if (k >= nblk) break;
        const int b = 32 * k + lane;
        const long long s = bs[b] + bs[nb + b];
        const int m = (int)bs[2 * nb + b];
        const long long sfx = suffix_scan(s, lane);
        const int ne = __reduce_add_sync(FULL, m), nt = __reduce_add_sync(FULL, m > 0 ? 1 : 0);
        if (lane == 0) { s_tot[k] = sfx; atomicAdd(&s_ne, ne); atomicAdd(&s_nt, nt); }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive suffix sums of the block totals
        long long carry = 0;
        for (int j = nsup - 1; j >= 0; --j) {
            const int k = 32 * j + lane;
            const long long v = k < nblk ? s_tot[k] : 0;
            const long long sfx = suffix_scan(v, lane);
            if (k < nblk) s_off[k] = carry + sfx - v;
            carry += __shfl_sync(FULL, sfx, 0);
        }
    }
    __syncthreads();
    for (int j = 0; j < nsup; ++j) {  // per-bin terms, level-1 scans
        const int k = warp + 32 * j;
        if (k >= nblk) break;
        const int b = 32 * k + lane;
        const long long e = bs[nb + b];
        const long long s = bs[b] + e;
        const int m = (int)bs[2 * nb + b];
        const long long sfx = suffix_scan(s, lane);
        const BinTerms bt = bin_terms(s_off[k] + sfx, e, m, efron ? 1 : 0);
        __syncwarp();
        const double gi = hs_scan(bt.G, lane), ti = hs_scan(bt.T, lane);
        sg[b] = gi; sf[b] = bt.F;
        if (lane == 31) { s_g[k] = gi; s_t[k] = ti; }
    }
    __syncthreads();
    unsigned char *seg_state = state + seg * seg_state_stride(nb);
    float2 *table = reinterpret_cast<float2 *>(seg_state + sizeof(b200surv_cox_header));
    double run_g = 0.0, run_t = 0.0;
    for (int j = 0; j < nsup; ++j) {  // levels 2 and 3 (every warp redundantly), table
        const int kk = 32 * j + lane;
        const double vg = kk < nblk ? s_g[kk] : 0.0, vt = kk < nblk ? s_t[kk] : 0.0;
        const double ig = hs_scan(vg, lane), it = hs_scan(vt, lane);
        const double ex = __shfl_sync(FULL, ig, warp > 0 ? warp - 1 : 0);
        const int k = warp + 32 * j;
        if (k < nblk) {
            const int b = 32 * k + lane;
            const double P = (run_g + (warp > 0 ? ex : 0.0)) + sg[b];
            table[b] = make_float2((float)P, (float)sf[b]);
        }
        const int lastpos = min(31, nblk - 1 - 32 * j);
        run_g += __shfl_sync(FULL, ig, lastpos);
        run_t += __shfl_sync(FULL, it, lastpos);
    }
    if (t == 0) {
        HeaderIn h;
        h.sum_ev_eta_q = bs[3 * (size_t)nb]; h.n_not_binnable = bs[3 * (size_t)nb + 1];
        h.sum_w_ceil = bs[3 * (size_t)nb + 2]; h.n_bad_time = bs[3 * (size_t)nb + 3];
        h.max_eta = bins_max[2 * seg]; h.min_eta = -bins_max[2 * seg + 1];
        h.n_events = s_ne; h.n_times = s_nt; h.T = run_t; h.peer_timeout = 0;
        write_header(h, efron ? 1 : 0, reduction, shift, nb, reinterpret_cast<b200surv_cox_header *>(seg_state),
                     out_loss + seg);
    }
}

// ================================================================ fused forward (one cohort, one launch)
// Pass 1, the exact reduction of the CTA partials and the whole O(nbins) tail in ONE cooperative kernel with a
// single grid barrier (after the partial histograms are flushed).  The bins are dealt out in blocks of 32 (one warp
// per block, one or two blocks per CTA on a 148-SM grid); the cross-block dependencies of the tail -- the suffix
// sums D and the prefix sums P -- travel through self-flagged 64-bit words in global memory (decoupled look-back:
// a block publishes its total, then reads the totals of the blocks after / before it), not through grid barriers.
struct TailSlot {
    unsigned long long A;  // block total of S (36.28 fixed point, < 2^62) | SLOT_FLAG
    unsigned long long C;  // events << 32 | distinct event times of the block | SLOT_FLAG
    unsigned long long G;  // block total of G (double bits); SLOT_EMPTY until published
    unsigned long long T;  // block total of T
};
constexpr unsigned long long SLOT_FLAG = 1ull << 63;
constexpr unsigned long long SLOT_EMPTY = ~0ull;  // a NaN pattern no arithmetic result carries
constexpr int SLOT_SPIN_MAX = 1 << 22;            // a bug must not hang the GPU: give up after ~seconds

__device__ __forceinline__ unsigned long long ld_slot(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_slot(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long wait_flagged(const unsigned long long *p) {
    unsigned long long v = ld_slot(p);
    for (int it = 0; !(v & SLOT_FLAG) && it < SLOT_SPIN_MAX; ++it) v = ld_slot(p);
    return v & ~SLOT_FLAG;
}
// Look-back with all of a lane's loads in flight together: word `w` of the slots k0 + lane + 32 j (j < LB_MAX) that
// lie below kend, each re-read until published (rare: the words are usually there when the reader arrives; issued one
// after the other the round trips added up to 3 us for the blocks that need all 128 words).
constexpr int LB_MAX = B200SURV_COX_MAX_BINS / 32 / 32;  // 8
template <bool FLAGGED>
__device__ __forceinline__ bool slot_ready(unsigned long long v) { return FLAGGED ? (v & SLOT_FLAG) != 0 : v != SLOT_EMPTY; }
// Control flow stays WARP-UNIFORM (predicated loads, a vote decides whether to go round again): a warp that diverges
// here keeps taking the slow path of every later shuffle (WARPSYNC.COLLECTIVE, ~250 ns each; measured).
template <bool FLAGGED>
__device__ __forceinline__ void lookback_words(const TailSlot *slots, int w, int k0, int kend, int lane,
                                               unsigned long long (&v)[LB_MAX]) {
    const int nj = (kend - k0 + 31) >> 5;  // uniform
#pragma unroll
    for (int j = 0; j < LB_MAX; ++j) {
        const int k = k0 + lane + 32 * j;
        v[j] = FLAGGED ? SLOT_FLAG : 0ull;
        if (j < nj && k < kend) v[j] = ld_slot(reinterpret_cast<const unsigned long long *>(slots + k) + w);
    }
    for (int it = 0; it < SLOT_SPIN_MAX; ++it) {
        bool pending = false;
#pragma unroll
        for (int j = 0; j < LB_MAX; ++j) {
            const int k = k0 + lane + 32 * j;
            if (j < nj && k < kend && !slot_ready<FLAGGED>(v[j])) {
                v[j] = ld_slot(reinterpret_cast<const unsigned long long *>(slots + k) + w);
                pending |= !slot_ready<FLAGGED>(v[j]);
            }
        }
        if (!__any_sync(FULL, pending)) break;
    }
#pragma unroll
    for (int j = 0; j < LB_MAX; ++j) {
        const int k = k0 + lane + 32 * j;
        if (FLAGGED) v[j] &= ~SLOT_FLAG;
        if (!(j < nj && k < kend)) v[j] = 0ull;
    }
}
__device__ __forceinline__ double wait_double(const unsigned long long *p) {
    unsigned long long v = ld_slot(p);
    for (int it = 0; v == SLOT_EMPTY && it < SLOT_SPIN_MAX; ++it) v = ld_slot(p);
    return __longlong_as_double((long long)v);
}

// ---- multi-GPU exchange through peer memory (NVLink / NVSwitch), fused into the cooperative forward.
// Every rank owns one "peer buffer" that all ranks of the box have mapped (symmetric allocation):
//   [ slot 0 | slot 1 ]   slot = epoch & 1;  each slot: for every source rank r < PEER_MAX_WORLD
//                         int64 words[3][nb] (S_cens, S_event, m of rank r's rows) and 8 scalar words
// PUSH, no barrier, no flag, no fence: in step k the warp that owns a 32-bin block stores its rank-local sums
// straight into every peer's buffer (region of source = this rank), each 64-bit word carrying the step's 2-bit tag
// (k & 3, never 0 on first use) in its top bits; then it polls the same block in its OWN buffer until the words of
// all sources show the tag, and adds them.  A word is its own "arrived" mark, so nothing has to be ordered against
// anything.  Integer sums: every rank obtains bit-identical totals.  Double buffering is enough: a rank rewrites
// slot k&1 in step k+2, after it has received every peer's step-k+1 words, which a peer only sends from its
// step-k+1 kernel, i.e. after its step-k kernel (the reader of this slot) has finished.
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_SCALARS = 8;  // sum of event log_hz, NOT_BINNABLE, ceil(sum w), BAD_TIME, max log_hz (float bits), 3 spare
constexpr unsigned long long PEER_VAL_MASK = (1ull << 62) - 1;
constexpr long long PEER_SPIN_LIMIT_NS = 2000000000ll;  // 2 s: a missing peer must not hang the GPU

__host__ __device__ inline size_t peer_src_words(int nb) { return 3 * (size_t)nb + PEER_SCALARS; }
__host__ __device__ inline size_t peer_slot_bytes(int nb) { return PEER_MAX_WORLD * peer_src_words(nb) * sizeof(long long); }

struct PeerArgs {
    int world, rank;
    unsigned epoch;
    int *status;  // local: set to 1 when the wait timed out
    unsigned char *buf[PEER_MAX_WORLD];
};

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long global_timer_ns() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}

struct FusedArgs {
    int ties, reduction;
    float *out_loss;
    unsigned char *state;
    TailSlot *slots;   // [nb / 32]
    long long *trace;  // optional (may be null): globaltimer stamps of the phases
    int nparts;        // CTAs [0, nparts) take part in pass 1 and own a partial histogram + CtaRec
};

struct TailArgs {
    TailSlot *slots;
    int nb, efron, reduction;
    float shift;
    float2 *table;
    b200surv_cox_header *hdr;
    float *out_loss;
    const CtaRec *recs;  // single GPU: the header scalars come from the CTA records
    int nparts;
    const PeerArgs *peer;  // multi GPU: from the scalar words of every peer's slot
    int peer_timeout;
    long long *trace;
};

// Multi-GPU exchange of `nw` (<= 4) words per lane: store them, tagged, into every peer's buffer at word offset `off`
// of this rank's source region, then poll the same words of every source in this rank's own buffer and add them up
// (this rank's own contribution comes from the registers).  Warp-uniform control flow; returns false on time-out.
__device__ __forceinline__ bool peer_push_sum(const PeerArgs &pa, int nb, const size_t (&off)[4], int nw, bool active,
                                              long long (&val)[4]) {
    const unsigned long long tag = (unsigned long long)(pa.epoch & 3u) << 62;
    const size_t slot_words = (size_t)(pa.epoch & 1u) * PEER_MAX_WORLD * peer_src_words(nb);
    const size_t src_words = peer_src_words(nb);
    if (active) {
        for (int p = 0; p < pa.world; ++p) {
            if (p == pa.rank) continue;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(pa.buf[p]) + slot_words + (size_t)pa.rank * src_words;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < nw) st_relaxed_sys_u64(dst + off[i], ((unsigned long long)val[i] & PEER_VAL_MASK) | tag);
        }
    }
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(pa.buf[pa.rank]) + slot_words;
    const long long t0 = global_timer_ns();
    bool ok = true;
    // sources in groups of four, the polls of a group in flight together (one after the other, seven sources cost seven
    // L2 round trips); uniform trip counts, a vote decides whether to go round again
    for (int p0 = 0; p0 < pa.world && ok; p0 += 4) {
        unsigned long long w[4][4];
        for (;;) {
            bool pending = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + q;
                const bool src_ok = p < pa.world && p != pa.rank;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[q][i] = tag;
                    if (src_ok && active && i < nw) {
                        w[q][i] = ld_relaxed_sys_u64(mine + (size_t)p * src_words + off[i]);
                        pending |= (w[q][i] >> 62) != (tag >> 62);
                    }
                }
            }
            if (!__any_sync(FULL, pending)) break;
            if (global_timer_ns() - t0 > PEER_SPIN_LIMIT_NS) { ok = false; break; }
        }
        ok = __all_sync(FULL, ok);
        if (!ok) break;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = p0 + q;
            if (p < pa.world && p != pa.rank && active) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nw) val[i] += (long long)(w[q][i] << 2) >> 2;  // 62-bit two's complement -> 64 bits
            }
        }
    }
    return ok;
}

// The tail of one 32-bin block, run by one warp (lane = bin).  s = S_cens + S_event, e = S_event (fixed point),
// m = event count of the lane's bin.  Data flow, no barrier: the block waits for the A words of the later blocks and
// for the G words of the earlier ones, each lane polling the words it needs.  (A __syncwarp() follows every section
// in which the lanes may drift apart -- polls, bin_terms: a shuffle on a diverged warp takes a ~100-cycle slow path
// per instruction, microseconds for a scan.)
__device__ __noinline__ void tail_block(const TailArgs &a, int blk, int lane, long long s, long long e, int m) {
    TailSlot *slots = a.slots;
    const int nblk = a.nb >> 5;
    const bool tr = a.trace != nullptr && blk == 0 && lane == 0;
#define BSTAMP(i) do { if (a.trace != nullptr) { if (lane == 0) a.trace[32 + 8 * blk + (i)] = global_timer_ns(); __syncwarp(); } } while (0)
    BSTAMP(0);
    // ---- D: suffix sums (integers).  Publish the block total, look back over the later blocks.
    const long long sfx = suffix_scan(s, lane);
    const unsigned ne = __reduce_add_sync(FULL, (unsigned)m), nt = __reduce_add_sync(FULL, m > 0 ? 1u : 0u);
    if (lane == 0) {
        st_slot(&slots[blk].A, (unsigned long long)sfx | SLOT_FLAG);
        st_slot(&slots[blk].C, ((unsigned long long)ne << 32) | nt | SLOT_FLAG);
    }
    unsigned long long v[LB_MAX];
    lookback_words<true>(slots, 0, blk + 1, nblk, lane, v);
    long long off = 0;
#pragma unroll
    for (int j = 0; j < LB_MAX; ++j) off += (long long)v[j];
    off = warp_sum(off);
    BSTAMP(1);
    if (tr) a.trace[7] = global_timer_ns();
    // ---- per-bin terms, level-1 scans, publish the block totals
    const BinTerms bt = bin_terms(off + sfx, e, m, a.efron);
    __syncwarp();
    const double gi = hs_scan(bt.G, lane), ti = hs_scan(bt.T, lane);
    if (lane == 31) {
        st_slot(&slots[blk].G, (unsigned long long)__double_as_longlong(gi));
        st_slot(&slots[blk].T, (unsigned long long)__double_as_longlong(ti));
    }
    BSTAMP(2);
    if (tr) a.trace[8] = global_timer_ns();
    // ---- P: look back over the earlier blocks (levels 2 and 3 of the canonical tree)
    const int sblk = blk >> 5, pos = blk & 31;
    lookback_words<false>(slots, 2, 0, blk, lane, v);  // G of the blocks [0, blk): lane l holds blocks l, l + 32, ...
    double run_g = 0.0, ex = 0.0;
#pragma unroll
    for (int j = 0; j < LB_MAX; ++j) {
        if (j <= sblk) {  // (zero from this block on: the inclusive scan at a lane only depends on the lanes before it)
            const double ig = hs_scan(__longlong_as_double((long long)v[j]), lane);
            if (j < sblk) run_g += __shfl_sync(FULL, ig, 31);
            else ex = __shfl_sync(FULL, ig, pos > 0 ? pos - 1 : 0);
        }
    }
    const double P = (run_g + (pos > 0 ? ex : 0.0)) + gi;
    a.table[32 * blk + lane] = make_float2((float)P, (float)bt.F);
    BSTAMP(3);
    if (tr) a.trace[9] = global_timer_ns();
#undef BSTAMP
}

// Loss and header, by a warp that owns no block: waits for the T totals and the event counts of all blocks, forms the
// total of T along the canonical tree, writes the header.
__device__ __noinline__ void tail_header(const TailArgs &a, int lane) {
    TailSlot *slots = a.slots;
    const int nblk = a.nb >> 5;
    HeaderIn h;
    {  // this rank's scalar words, from its CTA records
        double se = 0.0, sw = 0.0;
        float mx = -INFINITY, nmn = -INFINITY;
        unsigned fl = 0;
        for (int c = lane; c < a.nparts; c += 32) {
            const CtaRec *rp = a.recs + c;
            se += __ldcg(&rp->sum_ev_eta); sw += __ldcg(&rp->sum_w); mx = fmaxf(mx, __ldcg(&rp->max_eta));
            nmn = fmaxf(nmn, -__ldcg(&rp->min_eta));
            fl |= __ldcg(&rp->flags);
        }
        se = warp_sum(se); sw = warp_sum(sw); mx = warp_max(mx); nmn = warp_max(nmn); fl = warp_or(fl);
        h.sum_ev_eta_q = __double2ll_rn(se * ETA_SCALE);
        h.n_not_binnable = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1 : 0;
        h.sum_w_ceil = (long long)fmin(ceil(sw), 1.0e18);  // (fits the 62-bit words of the multi-GPU exchange)
        h.n_bad_time = (fl & B200SURV_COXF_BAD_TIME) ? 1 : 0;
        h.max_eta = mx; h.min_eta = -nmn;
    }
    int peer_timeout = a.peer_timeout;
    if (a.peer != nullptr) {  // add the other ranks' scalar words (lanes 0..3), maximum of max log_hz over the ranks
        long long val[4] = {0, 0, 0, 0};
        if (lane == 0) val[0] = h.sum_ev_eta_q;
        if (lane == 1) val[0] = h.n_not_binnable;
        if (lane == 2) val[0] = h.sum_w_ceil;
        if (lane == 3) val[0] = h.n_bad_time;
        const size_t off[4] = {3 * (size_t)a.nb + (size_t)(lane < 4 ? lane : 0), 0, 0, 0};
        if (!peer_push_sum(*a.peer, a.nb, off, 1, lane < 4, val)) peer_timeout = 1;
        h.sum_ev_eta_q = __shfl_sync(FULL, val[0], 0); h.n_not_binnable = __shfl_sync(FULL, val[0], 1);
        h.sum_w_ceil = __shfl_sync(FULL, val[0], 2); h.n_bad_time = __shfl_sync(FULL, val[0], 3);
        // the maximum: one word per source rank, lane p reads source p
        const unsigned long long tag = (unsigned long long)(a.peer->epoch & 3u) << 62;
        const size_t slot_words = (size_t)(a.peer->epoch & 1u) * PEER_MAX_WORLD * peer_src_words(a.nb);
        const size_t word = 3 * (size_t)a.nb + 4;
        if (lane < a.peer->world && lane != a.peer->rank)
            st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(a.peer->buf[lane]) + slot_words +
                                   (size_t)a.peer->rank * peer_src_words(a.nb) + word,
                               (unsigned long long)__float_as_uint(h.max_eta) |
                                   ((unsigned long long)(__float_as_uint(-h.min_eta) >> 2) << 32) | tag);
        // (-min travels with its two lowest mantissa bits dropped and is rounded towards a smaller minimum on arrival;
        // this rank's own value goes through the same rounding, so that every rank ends up with the same word)
        auto dec_nmn = [](unsigned field30) {
            const unsigned hb = field30 << 2;
            return __uint_as_float((hb & 0x80000000u) ? hb : (hb | 3u));
        };
        float mx = h.max_eta, nmn = dec_nmn(__float_as_uint(-h.min_eta) >> 2);
        const long long t0 = global_timer_ns();
        for (;;) {
            bool pending = false;
            if (lane < a.peer->world && lane != a.peer->rank) {
                const unsigned long long w = ld_relaxed_sys_u64(reinterpret_cast<const unsigned long long *>(a.peer->buf[a.peer->rank]) +
                                                                slot_words + (size_t)lane * peer_src_words(a.nb) + word);
                pending = (w >> 62) != (tag >> 62);
                if (!pending) {
                    mx = __uint_as_float((unsigned)w);
                    nmn = fmaxf(nmn, dec_nmn((unsigned)((w >> 32) & 0x3fffffffu)));
                }
            }
            if (!__any_sync(FULL, pending)) break;
            if (global_timer_ns() - t0 > PEER_SPIN_LIMIT_NS) { peer_timeout = 1; break; }
        }
        peer_timeout = __any_sync(FULL, peer_timeout != 0) ? 1 : 0;
        h.max_eta = peer_timeout ? 0.f : warp_max(mx);
        h.min_eta = peer_timeout ? 0.f : -warp_max(nmn);
        if (peer_timeout) { h.sum_ev_eta_q = 0; h.n_not_binnable = 0; h.sum_w_ceil = 0; h.n_bad_time = 0; }
    }
    unsigned long long vc[LB_MAX], vt[LB_MAX];
    lookback_words<true>(slots, 1, 0, nblk, lane, vc);
    lookback_words<false>(slots, 3, 0, nblk, lane, vt);
    long long cnt_e = 0, cnt_t = 0;
    double run_t = 0.0;
#pragma unroll
    for (int j = 0; j < LB_MAX; ++j) {
        cnt_e += (long long)(vc[j] >> 32); cnt_t += (long long)(vc[j] & 0xffffffffull);
        if (32 * j < nblk)
            run_t += __shfl_sync(FULL, hs_scan(__longlong_as_double((long long)vt[j]), lane), min(31, nblk - 1 - 32 * j));
    }
    cnt_e = warp_sum(cnt_e); cnt_t = warp_sum(cnt_t);
    h.n_events = cnt_e; h.n_times = cnt_t; h.T = run_t;
    h.peer_timeout = (peer_timeout || (a.peer != nullptr && __ldcg(a.peer->status) != 0)) ? 1 : 0;
    if (lane == 0) {
        write_header(h, a.efron, a.reduction, a.shift, a.nb, a.hdr, a.out_loss);
        if (a.trace != nullptr) a.trace[10] = global_timer_ns();
    }
}

template <bool PEER>
__global__ void __launch_bounds__(P1_THREADS, 1)
cox_binned_fwd_fused(const float *__restrict__ log_hz, const float *__restrict__ time,
                     const uint8_t *__restrict__ event, int64_t n, int nb, float shift, int vec_ok,
                     unsigned char *__restrict__ partial, CtaRec *__restrict__ recs, const FusedArgs fa, int p1_mode,
                     const __grid_constant__ PeerArgs pa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int cta = blockIdx.x, G = gridDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nblk = nb >> 5, bpc = (nblk + G - 1) / G;  // 32-bin blocks per CTA (host: bpc <= 32)
    const int blk0 = cta * bpc, nparts = fa.nparts;
#define TRACE(i)                                                                          \
    do {                                                                                  \
        if (fa.trace != nullptr && cta == 0 && t == 0) fa.trace[i] = global_timer_ns();  \
    } while (0)
    TRACE(0);
    if (t < bpc && blk0 + t < nblk) {  // look-back slots of this CTA's blocks (ordered by the grid barrier)
        TailSlot *sl = fa.slots + blk0 + t;
        sl->A = 0; sl->C = 0; sl->G = SLOT_EMPTY; sl->T = SLOT_EMPTY;
    }
    if (PEER && cta == 0 && t == 0) *pa.status = 0;  // set again if a peer never arrives
    if (cta < nparts) {
        if (p1_mode == 2)
            pass1_body_ring(log_hz, time, event, n, nb, shift, partial, recs, cta, nparts, smem_raw);
        else if (p1_mode == 1)
            pass1_body_tma(log_hz, time, event, n, nb, shift, partial, recs, cta, nparts, smem_raw);
        else
            pass1_body(log_hz, time, event, nullptr, n, nb, shift, vec_ok, partial, recs, 0, cta, nparts,
                       reinterpret_cast<unsigned *>(smem_raw));
    }
    TRACE(1);
    grid.sync();
    TRACE(2);

    // ---- exact reduction of the partials for this CTA's blocks: lanes along the 32 bins of a block (contiguous
    // 256-byte reads of every partial record), warps along the records, then a cross-warp sum through shared
    // memory (the histogram storage is free again); warp j keeps the sums of block blk0 + j
    long long my_c = 0, my_e = 0;
    int my_m = 0;
    {
        unsigned long long *s_c = reinterpret_cast<unsigned long long *>(smem_raw);  // [32 warps][32 bins]
        unsigned long long *s_e = s_c + 1024;
        unsigned *s_mm = reinterpret_cast<unsigned *>(s_e + 1024);
        for (int j = 0; j < bpc && blk0 + j < nblk; ++j) {
            const int b = 32 * (blk0 + j) + lane;
            unsigned long long vc[RED_MAX_ITERS], ve[RED_MAX_ITERS];
            unsigned vm[RED_MAX_ITERS];
#pragma unroll
            for (int k = 0; k < RED_MAX_ITERS; ++k) {
                const int c = warp + 32 * k;
                const bool in = c < nparts;
                const unsigned char *pc = partial + (size_t)(in ? c : 0) * PARTIAL_BYTES_PER_BIN * (size_t)nb;
                vc[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + b) : 0ull;
                ve[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + nb + b) : 0ull;
                vm[k] = in ? __ldcg(reinterpret_cast<const unsigned *>(pc + 16 * (size_t)nb) + b) : 0u;
            }
            unsigned long long sc = 0, se = 0;
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < RED_MAX_ITERS; ++k) { sc += vc[k]; se += ve[k]; m += vm[k]; }
            if (j) __syncthreads();  // the previous block's sums have been read
            s_c[t] = sc; s_e[t] = se; s_mm[t] = m;
            __syncthreads();
            if (warp == j) {
                unsigned long long c0 = 0, c1 = 0, e0 = 0, e1 = 0;
                unsigned m0 = 0, m1 = 0;
#pragma unroll 8
                for (int w = 0; w < 32; w += 2) {
                    c0 += s_c[w * 32 + lane]; c1 += s_c[w * 32 + 32 + lane];
                    e0 += s_e[w * 32 + lane]; e1 += s_e[w * 32 + 32 + lane];
                    m0 += s_mm[w * 32 + lane]; m1 += s_mm[w * 32 + 32 + lane];
                }
                my_c = (long long)(c0 + c1); my_e = (long long)(e0 + e1); my_m = (int)(m0 + m1);
            }
        }
    }
    TRACE(3);
    const bool has_blk = warp < bpc && blk0 + warp < nblk;
    const int blk = blk0 + warp;
    int peer_timeout = 0;
    if constexpr (PEER) {
        // ---- push this block's rank-local sums to every peer, wait for theirs, add (no barrier: see PeerArgs)
        if (has_blk) {
            const size_t b = 32 * (size_t)blk + lane;
            const size_t off[4] = {b, (size_t)nb + b, 2 * (size_t)nb + b, 0};
            long long val[4] = {my_c, my_e, (long long)my_m, 0};
            if (!peer_push_sum(pa, nb, off, 3, true, val)) {  // a peer never arrived: leave an empty cohort behind
                peer_timeout = 1;
                val[0] = 0; val[1] = 0; val[2] = 0;
                if (lane == 0) atomicExch(pa.status, 1);
            }
            my_c = val[0]; my_e = val[1]; my_m = (int)val[2];
        }
        TRACE(6);
    }

    // ---- the O(nbins) tail: one warp per block (the first warps of the CTA), the header by the last warp of the CTA
    // that owns the last block; no further barrier
    const bool is_hdr = cta == (nblk - 1) / bpc && warp == P1_THREADS / 32 - 1;
    if (has_blk || is_hdr) {
        TailArgs ta;
        ta.slots = fa.slots; ta.nb = nb; ta.efron = fa.ties == B200SURV_TIES_EFRON ? 1 : 0; ta.reduction = fa.reduction;
        ta.shift = shift;
        ta.table = reinterpret_cast<float2 *>(fa.state + sizeof(b200surv_cox_header));
        ta.hdr = reinterpret_cast<b200surv_cox_header *>(fa.state);
        ta.out_loss = fa.out_loss;
        ta.recs = recs; ta.nparts = nparts;
        ta.peer = PEER ? &pa : nullptr; ta.peer_timeout = peer_timeout;
        ta.trace = fa.trace;
        if (has_blk) tail_block(ta, blk, lane, my_c + my_e, my_e, my_m);
        else tail_header(ta, lane);
    }
#undef TRACE
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
    // rolling prefetch: the loads of the next two groups are in flight while the current two are computed
    {
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), t0 = e0, e1 = e0, t1 = e0;
        uint32_t v0 = 0, v1 = 0;
        // REVERSE traversal (virtual index q counts down from the last group): the rows pass 1 read last are
        // still L2-resident, so the backward pass starts there
        const int64_t last = ngroups - 1;
        bool h0 = g < ngroups, h1 = g + stride < ngroups;
        if (h0) { e0 = ldg_stream_f4(lh + 4 * (last - g)); t0 = ldg_stream_f4(tm + 4 * (last - g)); v0 = ldg_stream_u32(evp + 4 * (last - g)); }
        if (h1) { e1 = ldg_stream_f4(lh + 4 * (last - g - stride)); t1 = ldg_stream_f4(tm + 4 * (last - g - stride)); v1 = ldg_stream_u32(evp + 4 * (last - g - stride)); }
        while (h0) {
            const int64_t gn = g + 2 * stride;
            const bool n0 = gn < ngroups, n1 = gn + stride < ngroups;
            float4 ne0 = e0, nt0 = t0, ne1 = e1, nt1 = t1;
            uint32_t nv0 = 0, nv1 = 0;
            if (n0) { ne0 = ldg_stream_f4(lh + 4 * (last - gn)); nt0 = ldg_stream_f4(tm + 4 * (last - gn)); nv0 = ldg_stream_u32(evp + 4 * (last - gn)); }
            if (n1) { ne1 = ldg_stream_f4(lh + 4 * (last - gn - stride)); nt1 = ldg_stream_f4(tm + 4 * (last - gn - stride)); nv1 = ldg_stream_u32(evp + 4 * (last - gn - stride)); }
            float4 o0;
            o0.x = p2_row(e0.x, t0.x, (v0 & 0xffu) != 0, c2, k, tab, nb);
            o0.y = p2_row(e0.y, t0.y, (v0 & 0xff00u) != 0, c2, k, tab, nb);
            o0.z = p2_row(e0.z, t0.z, (v0 & 0xff0000u) != 0, c2, k, tab, nb);
            o0.w = p2_row(e0.w, t0.w, (v0 & 0xff000000u) != 0, c2, k, tab, nb);
            stg_stream_f4(og + 4 * (last - g), o0);
            if (h1) {
                float4 o1;
                o1.x = p2_row(e1.x, t1.x, (v1 & 0xffu) != 0, c2, k, tab, nb);
                o1.y = p2_row(e1.y, t1.y, (v1 & 0xff00u) != 0, c2, k, tab, nb);
                o1.z = p2_row(e1.z, t1.z, (v1 & 0xff0000u) != 0, c2, k, tab, nb);
                o1.w = p2_row(e1.w, t1.w, (v1 & 0xff000000u) != 0, c2, k, tab, nb);
                stg_stream_f4(og + 4 * (last - g - stride), o1);
            }
            e0 = ne0; t0 = nt0; v0 = nv0; e1 = ne1; t1 = nt1; v1 = nv1;
            h0 = n0; h1 = n1; g = gn;
        }
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
    size_t off_status, off_trace, off_slots, off_partial, off_recs, off_bins, off_bins_max, off_scr_g, off_scr_f, total;
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
    L.off_status = take(sizeof(int));
    L.off_trace = take((32 + 8 * 256) * sizeof(long long));
    L.off_slots = take((size_t)(nb / 32) * sizeof(TailSlot));
    L.off_partial = take((size_t)n_seg * L.nctas * PARTIAL_BYTES_PER_BIN * nb);
    L.off_recs = take((size_t)n_seg * L.nctas * sizeof(CtaRec));
    L.off_bins = take((size_t)n_seg * (3 * (size_t)nb + 4) * sizeof(long long));
    L.off_bins_max = take((size_t)n_seg * 2 * sizeof(float));
    L.off_scr_g = take((size_t)n_seg * nb * sizeof(double));
    L.off_scr_f = take((size_t)n_seg * nb * sizeof(double));
    L.total = o;
    return L;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool coop_supported() {  // of the current device (not cached: one attribute query, microseconds)
    const int dev = current_device();
    int v = 0;
    return dev >= 0 && cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v &&
           num_sms() >= 16;  // (the fused kernel wants bpc < 32 blocks of bins per CTA)
}

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
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(B200SURV_COX_MAX_BINS * PARTIAL_BYTES_PER_BIN)));
        attr_once.mark();
    }
    unsigned char *partial = w8 + L.off_partial;
    CtaRec *recs = reinterpret_cast<CtaRec *>(w8 + L.off_recs);
    cox_binned_pass1<<<dim3(L.nctas, (unsigned)n_seg), P1_THREADS, smem, st>>>(
        log_hz, time, event, seg_off, n, nb, shift, vec_ok, partial, recs);
    cox_binned_reduce<<<dim3(nb / RED_BINS, (unsigned)n_seg), RED_THREADS, 0, st>>>(partial, recs, L.nctas, nb, bins,
                                                                                    bins_max);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(2);
    return B200SURV_OK;
}

int32_t launch_finish(const long long *bins, const float *bins_max, int64_t n_seg, int ties, int reduction, int nb,
                      float shift, float *out_loss, void *state, const BinnedLayout &L, unsigned char *w8,
                      cudaStream_t st) {
    cox_binned_finish<<<(unsigned)n_seg, IT_THREADS, 0, st>>>(
        bins, bins_max, nb, ties, reduction, shift, reinterpret_cast<double *>(w8 + L.off_scr_g),
        reinterpret_cast<double *>(w8 + L.off_scr_f), out_loss, static_cast<unsigned char *>(state));
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(1);
    return B200SURV_OK;
}

// one cooperative launch: pass 1 + reduce (+ peer exchange) + the O(nbins) tail
int32_t launch_fused(const float *log_hz, const float *time, const uint8_t *event, int64_t n, int ties, int reduction,
                     int nb, float shift, float *out_loss, void *state, const BinnedLayout &L, unsigned char *w8,
                     const PeerArgs *peer, cudaStream_t st) {
    int vec_ok = aligned16(log_hz) && aligned16(time) && ((reinterpret_cast<uintptr_t>(event) & 3) == 0);
    // Pass-1 staging, B200SURV_P1 = ring (default) | reg | tma.  Measured at 16.7M rows, forward only:
    // per-thread cp.async ring 64.4 us, register-staged loads 70.3 us, TMA bulk ring + producer warp 78.9 us
    // (profiles/r1_v12_p1_staging_ab.txt, r1_v10_tma_ab.txt).
    static const int p1_env = [] {
        const char *e = getenv("B200SURV_P1");
        if (e && !strcmp(e, "tma")) return 1;
        if (e && !strcmp(e, "reg")) return 0;
        return 2;
    }();
    int p1_mode = (vec_ok && nb <= 4096 && n >= 4 * (int64_t)TMA_TILE) ? p1_env : 0;
    size_t smem = p1_mode == 2 ? ring_smem_bytes(nb) : p1_mode == 1 ? tma_smem_bytes(nb) : (size_t)nb * 24 + 16;
    if (smem < 20480) smem = 20480;  // the reduce step stages 32 x 32 x (8 + 8 + 4) bytes
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        size_t mx = (size_t)B200SURV_COX_MAX_BINS * 24 + 16;
        if (tma_smem_bytes(4096) > mx) mx = tma_smem_bytes(4096);
        if (ring_smem_bytes(4096) > mx) mx = ring_smem_bytes(4096);
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
        attr_once.mark();
    }
    unsigned char *partial = w8 + L.off_partial;
    CtaRec *recs = reinterpret_cast<CtaRec *>(w8 + L.off_recs);
    static const bool trace_on = getenv("B200SURV_PEER_TRACE") != nullptr;
    FusedArgs fa;
    fa.ties = ties; fa.reduction = reduction; fa.out_loss = out_loss; fa.state = static_cast<unsigned char *>(state);
    fa.slots = reinterpret_cast<TailSlot *>(w8 + L.off_slots);
    fa.trace = trace_on ? reinterpret_cast<long long *>(w8 + L.off_trace) : nullptr;
    fa.nparts = L.nctas;
    int nb_i = nb;
    PeerArgs pa;
    if (peer) pa = *peer; else memset(&pa, 0, sizeof(pa));
    void *args[] = {(void *)&log_hz, (void *)&time, (void *)&event, (void *)&n, (void *)&nb_i, (void *)&shift,
                    (void *)&vec_ok, (void *)&partial, (void *)&recs, (void *)&fa, (void *)&p1_mode, (void *)&pa};
    const void *fn = peer ? (const void *)cox_binned_fwd_fused<true> : (const void *)cox_binned_fwd_fused<false>;
    // the full grid always: CTAs beyond the pass-1 participants still own blocks of bins in the reduce step and the tail
    B200_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(P1_THREADS), args, smem, st));
    count_launches(1);
    return B200SURV_OK;
}

}  // namespace

// ================================================================ internal entry points
size_t cox_binned_state_bytes(int64_t n_seg, int nb) { return (size_t)n_seg * seg_state_stride(nb); }
size_t cox_binned_workspace_bytes(int64_t n, int64_t n_seg, int nb) { return binned_layout(n, n_seg, nb).total; }

int32_t cox_binned_partial(const float *log_hz, const float *time, const uint8_t *event,
                           const int64_t *seg_off, int64_t n, int64_t n_seg, int nb, float shift,
                           int64_t *bins_sum, float *bins_max, void *ws, size_t ws_bytes, cudaStream_t st) {
    int32_t rc = check_common(n, n_seg, nb);
    if (rc) return rc;
    const BinnedLayout L = binned_layout(n, n_seg, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    return launch_pass1_reduce(log_hz, time, event, seg_off, n, n_seg, nb, shift,
                               reinterpret_cast<long long *>(bins_sum), bins_max, L, static_cast<unsigned char *>(ws), st);
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
    return launch_finish(reinterpret_cast<const long long *>(bins_sum), bins_max, n_seg, ties, reduction, nb, shift,
                         out_loss, state, L, static_cast<unsigned char *>(ws), st);
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
    static const bool no_fuse = getenv("B200SURV_NO_FUSE") != nullptr;  // diagnostics: force the three-kernel path
    if (n_seg == 1 && seg_off == nullptr && coop_supported() && !no_fuse)
        return launch_fused(log_hz, time, event, n, ties, reduction, nb, shift, out_loss, state, L, w8, nullptr, st);
    rc = launch_pass1_reduce(log_hz, time, event, seg_off, n, n_seg, nb, shift, bins, bins_max, L, w8, st);
    if (rc) return rc;
    return launch_finish(bins, bins_max, n_seg, ties, reduction, nb, shift, out_loss, state, L, w8, st);
}

size_t cox_binned_peer_buffer_bytes(int nb) { return 2 * peer_slot_bytes(nb); }
size_t cox_binned_peer_trace_offset(int64_t n, int nb) { return binned_layout(n, 1, nb).off_trace; }

int32_t cox_binned_fwd_peer(const float *log_hz, const float *time, const uint8_t *event, int64_t n, int ties,
                            int reduction, int nb, float shift, float *out_loss, void *state, size_t state_bytes,
                            void *ws, size_t ws_bytes, void *const *peer_bufs, int world, int rank, unsigned epoch,
                            cudaStream_t st) {
    int32_t rc = check_common(n, 1, nb);
    if (rc) return rc;
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    B200_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "world in [1,16], rank in [0,world)");
    B200_REQUIRE(peer_bufs != nullptr && epoch != 0, "peer_bufs must hold `world` pointers and epoch starts at 1");
    if (!coop_supported()) { set_error("cox binned peer exchange needs cooperative launch"); return B200SURV_UNSUPPORTED; }
    const BinnedLayout L = binned_layout(n, 1, nb);
    if (ws_bytes < L.total) { set_error("cox binned: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    if (state_bytes < cox_binned_state_bytes(1, nb)) { set_error("cox binned: state buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world; pa.rank = rank; pa.epoch = epoch;
    pa.status = reinterpret_cast<int *>(w8 + L.off_status);
    for (int p = 0; p < world; ++p) {
        B200_REQUIRE(peer_bufs[p] != nullptr && (reinterpret_cast<uintptr_t>(peer_bufs[p]) & 255) == 0, "peer buffer alignment (256)");
        pa.buf[p] = static_cast<unsigned char *>(peer_bufs[p]);
    }
    return launch_fused(log_hz, time, event, n, ties, reduction, nb, shift, out_loss, state, L, w8, &pa, st);
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
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             B200SURV_COX_MAX_BINS * (int)sizeof(float2)));
        attr_once.mark();
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
    count_launches(1);
    return B200SURV_OK;
}

}  // namespace b200surv
