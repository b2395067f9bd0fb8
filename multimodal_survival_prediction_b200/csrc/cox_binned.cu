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
//   K2 reduce  per-CTA partials -> per-bin int64 sums (this is what multi-GPU exchanges); the
//              last CTA to finish suffix-scans them (fp64): D[b] = sum_{b' >= b} S[b'], E/(D m).
//   K3 items   Efron: log(x), 1/x, (l/m)/x summed over l < m, x = 1 - (l/m) E/D.  The (bin, l) pairs of
//              all bins are cut into 256-term chunks dealt out evenly to the warps (efron_chunks), each
//              chunk sum converted to 2^-27 fixed point and added with integer atomics: balanced for any
//              tie structure, exact, order independent.  The last CTA to finish forms
//              P[b] = sum_{b' <= b} G[b'], the loss, the header and the (P,F) table.
//   K4 pass 2  (backward) streams the rows again, 9 B read + 4 B write:
//              grad = scale * (d - w * (P[b] - d * F[b]))
// One cohort per call (the headline case) runs K1..K3 as ONE cooperative kernel, cox_binned_fwd_fused
// (grid barriers between the phases); with peers it also carries the multi-GPU exchange of the per-bin
// sums over NVLink peer memory (PeerArgs).  Every path produces bit-identical results.
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
constexpr int RED_THREADS = 1024;  // reduce: 32 bins x 32 groups of partials; also the scan block
constexpr int RED_BINS = 32;
constexpr int RED_NG = RED_THREADS / RED_BINS;
constexpr int RED_MAX_ITERS = 5;   // ceil(max pass-1 CTAs per segment / RED_NG): up to 160 CTAs
constexpr int MAX_P1_CTAS = RED_NG * RED_MAX_ITERS;
constexpr int IT_THREADS = 1024;   // items / finish
constexpr int EF_CHUNK = 256;                    // fused forward: granule of the Efron work split
constexpr double EF_SCALE = 134217728.0;         // 2^27 fixed point of the per-bin Efron sums (|sum| < 2^36)
constexpr double EF_INV = 1.0 / 134217728.0;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int FIX_BITS = 28;
constexpr double FIX_INV = 1.0 / 268435456.0;    // 2^-28
constexpr double ETA_SCALE = 16777216.0;         // 2^24: fixed point of the sum of event log_hz
constexpr double ETA_INV = 1.0 / 16777216.0;
// shift is suitable when -8 <= max(log_hz) - shift <= 20 and the sum of weights stays below 2^30
constexpr float SHIFT_HI = 20.f, SHIFT_LO = -8.f;
constexpr double SUMW_LIMIT = 1073741824.0;      // 2^30

struct CtaRec {
    double sum_ev_eta;  // sum of log_hz over this CTA's event rows
    double sum_w;       // sum of weights (overflow guard)
    float max_eta;
    unsigned flags;
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
    float se, mx, sw;
    bool notbin, badt;
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
    acc.sw += wq;
    acc.se += ev ? eta : 0.f;
    int bin = __float2int_rz(t);
    const bool ok = ((unsigned)bin < nb) && ((float)bin == t);
    acc.notbin |= !ok;
    acc.badt |= !(t >= 0.f);
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
// four rows of one 128-bit group: the four returning atomics are issued back to back, then the carries
__device__ __forceinline__ void p1_group(const float4 e, const float4 t, uint32_t v, float c2, unsigned nb,
                                         uint32_t h_addr, P1Acc &acc) {
    const P1Row r0 = p1_prep(e.x, t.x, (v & 0xffu) != 0, c2, nb, h_addr, acc);
    const P1Row r1 = p1_prep(e.y, t.y, (v & 0xff00u) != 0, c2, nb, h_addr, acc);
    const P1Row r2 = p1_prep(e.z, t.z, (v & 0xff0000u) != 0, c2, nb, h_addr, acc);
    const P1Row r3 = p1_prep(e.w, t.w, (v & 0xff000000u) != 0, c2, nb, h_addr, acc);
    const unsigned o0 = p1_atom_lo(r0), o1 = p1_atom_lo(r1), o2 = p1_atom_lo(r2), o3 = p1_atom_lo(r3);
    p1_finish(r0, o0); p1_finish(r1, o1); p1_finish(r2, o2); p1_finish(r3, o3);
}
__device__ __forceinline__ void p1_row(float eta, float t, bool ev, float c2, unsigned nb, uint32_t h_addr,
                                       P1Acc &acc) {
    const P1Row r = p1_prep(eta, t, ev, c2, nb, h_addr, acc);
    p1_finish(r, p1_atom_lo(r));
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
    const unsigned fl = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
    if (threadIdx.x == 0) {
        CtaRec rec;
        rec.sum_ev_eta = se; rec.sum_w = sw; rec.max_eta = mx; rec.flags = fl;
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
    P1Acc acc{0.f, -INFINITY, 0.f, false, false};
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
    P1Acc acc{0.f, -INFINITY, 0.f, false, false};
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
    P1Acc acc{0.f, -INFINITY, 0.f, false, false};
    double se_d = 0.0, sw_d = 0.0;
    const int64_t ngroups = n >> 2, stride = (int64_t)nctas * P1_THREADS;
    const int64_t g0 = (int64_t)cta * P1_THREADS + t;
    const int64_t iters = g0 < ngroups ? (ngroups - g0 + stride - 1) / stride : 0;
    auto issue = [&](int64_t i, int s) {
        if (i < iters) {
            const int64_t g = g0 + i * stride;
            const uint32_t so = (uint32_t)s * RING_STAGE_BYTES;
            cp_async_16(my_e + so, log_hz + 4 * g);
            cp_async_16(my_t + so, time + 4 * g);
            cp_async_4(my_v + so, event + 4 * g);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < RING_STAGES - 1; ++s) issue(s, s);
    int s = 0;
    for (int64_t i = 0; i < iters; ++i) {
        int sn = s + RING_STAGES - 1;
        if (sn >= RING_STAGES) sn -= RING_STAGES;
        issue(i + RING_STAGES - 1, sn);
        cp_async_wait<RING_STAGES - 1>();
        const uint32_t so = (uint32_t)s * RING_STAGE_BYTES;
        float4 e4, t4;
        uint32_t v4;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e4.x), "=f"(e4.y), "=f"(e4.z), "=f"(e4.w) : "r"(my_e + so));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w) : "r"(my_t + so));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v4) : "r"(my_v + so));
        p1_group(e4, t4, v4, c2, nbu, h_addr, acc);
        se_d += (double)acc.se; sw_d += (double)acc.sw; acc.se = 0.f; acc.sw = 0.f;
        if (++s == RING_STAGES) s = 0;
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
                 CtaRec *__restrict__ recs, unsigned *__restrict__ tickets_k2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (blockIdx.x == 0 && threadIdx.x == 0) tickets_k2[blockIdx.y] = 0;  // "last CTA done" ticket of the reduce kernel
    pass1_body(log_hz, time, event, seg_off, n, nb, shift, vec_ok, partial, recs, blockIdx.y, blockIdx.x, gridDim.x,
               reinterpret_cast<unsigned *>(smem_raw));
}

// ================================================================ block scans (1024 threads)
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

// The block-wide part of the per-bin scan in one pass (three barriers): exclusive scans of a double and an int over
// the threads, and the block totals of two more ints.  shd[33], shi[36].
struct BinScan {
    double run;  // exclusive prefix of loc
    int crun;    // exclusive prefix of locc
    int n_chunks, n_events, n_times;
};
__device__ __forceinline__ BinScan block_bin_scan(double loc, int locc, int locm, int net, double *shd, int *shi) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double incd = loc;
    int incc = locc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double ud = __shfl_up_sync(FULL, incd, o);
        const int uc = __shfl_up_sync(FULL, incc, o);
        if (lane >= o) { incd += ud; incc += uc; }
    }
    const int wm = __reduce_add_sync(FULL, locm), wt = __reduce_add_sync(FULL, net);
    if (threadIdx.x == 0) { shi[33] = 0; shi[34] = 0; }
    __syncthreads();
    if (lane == 31) { shd[wid] = incd; shi[wid] = incc; }
    if (lane == 0) { atomicAdd(&shi[33], wm); atomicAdd(&shi[34], wt); }
    __syncthreads();
    if (wid == 0) {
        const double wd = (lane < nw) ? shd[lane] : 0.0;
        const int wc = (lane < nw) ? shi[lane] : 0;
        double id = wd;
        int ic = wc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ud = __shfl_up_sync(FULL, id, o);
            const int uc = __shfl_up_sync(FULL, ic, o);
            if (lane >= o) { id += ud; ic += uc; }
        }
        shd[lane] = id - wd;  // exclusive warp offsets
        shi[lane] = ic - wc;
        if (lane == 31) { shd[32] = id; shi[32] = ic; }
    }
    __syncthreads();
    BinScan r;
    r.run = shd[wid] + (incd - loc);
    r.crun = shi[wid] + (incc - locc);
    r.n_chunks = shi[32]; r.n_events = shi[33]; r.n_times = shi[34];
    return r;
}

// The block-wide part of the finish in one pass (three barriers): exclusive scan of g over the threads and the block
// sum of tv (fixed butterfly order: deterministic).  shd[33], shd2[33].
__device__ __forceinline__ double block_exscan_and_sum(double g, double tv, double *shd, double *shd2, double *tsum) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double inc = g;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    const double wt = warp_sum(tv);
    __syncthreads();
    if (lane == 31) shd[wid] = inc;
    if (lane == 0) shd2[wid] = wt;
    __syncthreads();
    if (wid == 0) {
        const double w = (lane < nw) ? shd[lane] : 0.0;
        double winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += u;
        }
        shd[lane] = winc - w;
        const double ts = warp_sum((lane < nw) ? shd2[lane] : 0.0);
        if (lane == 0) shd2[32] = ts;
    }
    __syncthreads();
    *tsum = shd2[32];
    return shd[wid] + (inc - g);
}

// per-segment scratch in the workspace, produced by the scan and consumed by K3
struct ScanBufs {
    double *D;          // [nb] risk-set sums
    double *rm;         // [nb] E / (D m)  (0 where m == 0)
    double *tgf;        // [3][nb] Efron T, G, F per bin
    long long *totals;  // [2] n_events, n_event_times
    int *cp;            // [nb + 1] exclusive prefix of the Efron chunk counts in scan order (see efron_chunks)
    unsigned *ticket;   // "last CTA done" counter of K3
};
struct ScanBase {
    double *D, *rm, *tgf;
    long long *totals;
    int *big;
    unsigned *tickets_k2, *tickets_k3;
};
__device__ __forceinline__ ScanBufs scan_bufs(const ScanBase &w, int seg, int nb) {
    ScanBufs o;
    o.D = w.D + (size_t)seg * nb; o.rm = w.rm + (size_t)seg * nb; o.tgf = w.tgf + (size_t)seg * 3 * nb;
    o.totals = w.totals + 2 * (size_t)seg; o.cp = w.big + (size_t)seg * (nb + 1); o.ticket = w.tickets_k3 + seg;
    return o;
}

// bins layout per segment (int64): S_cens_q[nb], S_event_q[nb], m[nb], sum_ev_eta_q (2^-24 fixed point),
// n_not_binnable, ceil(sum_w), n_bad_time.
// One block of 1024 threads: suffix sums D, E/(D m), totals, list of big bins.  `bs` may have been written by
// other CTAs of the same launch (fused path): it is read through L2.
template <int MAXPER>
__device__ void scan_segment(const long long *bs, int nb, ScanBufs o, double *shd, int *shi) {
    const int t = threadIdx.x;
    const int per = nb / RED_THREADS > 0 ? nb / RED_THREADS : 1;  // threads beyond nb idle
    const int hi_b = nb - t * per;  // reversed chunk [hi_b - per, hi_b)
    long long rc[MAXPER], re[MAXPER], rmv[MAXPER];
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {  // all loads in flight together
        const int b = hi_b - 1 - k;
        const bool in = (k < per) && (b >= 0);
        rc[k] = in ? __ldcg(bs + b) : 0ll;
        re[k] = in ? __ldcg(bs + nb + b) : 0ll;
        rmv[k] = in ? __ldcg(bs + 2 * nb + b) : 0ll;
    }
    if (t == 0) *o.ticket = 0;
    double sv[MAXPER];
    double loc = 0.0;
    int locm = 0, net = 0, locc = 0;
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        sv[k] = ((double)(unsigned long long)rc[k] + (double)(unsigned long long)re[k]) * FIX_INV;
        loc += sv[k];
        locm += (int)rmv[k];
        net += rmv[k] > 0 ? 1 : 0;
        locc += ((int)rmv[k] + EF_CHUNK - 1) / EF_CHUNK;
    }
    const BinScan sc = block_bin_scan(loc, locc, locm, net, shd, shi);  // run: sum over all later bins
    double run = sc.run;
    int crun = sc.crun;
    const int totm = sc.n_events, totn = sc.n_times, totc = sc.n_chunks;
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        const int b = hi_b - 1 - k;
        if (k < per && b >= 0) {
            run += sv[k];
            const int m = (int)rmv[k];
            o.D[b] = run;
            o.rm[b] = m > 0 ? ((double)(unsigned long long)re[k] * FIX_INV) / (run * (double)m) : 0.0;
            o.cp[t * per + k] = crun;  // scan order p = nb-1-b
            crun += (m + EF_CHUNK - 1) / EF_CHUNK;
            o.tgf[b] = 0.0; o.tgf[nb + b] = 0.0; o.tgf[2 * nb + b] = 0.0;  // Efron accumulators (all-zero bits)
        }
    }
    if (t == 0) { o.totals[0] = totm; o.totals[1] = totn; o.cp[nb] = totc; }
}

// Efron terms of one cohort, shared by the fused forward and the K3 kernel: the (bin, l) pairs of ALL bins, cut
// into chunks of EF_CHUNK consecutive l of one bin, are dealt out in equal contiguous runs to the W warps taking
// part -- balanced whatever the tie structure (one bin holding every event, or thousands of moderately tied
// ones; the work is the number of events).  A warp sums its run bin by bin (fp32, fp64 where x < 0.5) and adds the
// three sums to per-bin 2^-27 fixed-point integers: exact and order independent, so the loss is bit-reproducible
// and identical on every code path.  cp[p], p = nb-1-b: exclusive prefix of the chunk counts, cp[nb] = total.
__device__ __forceinline__ void efron_terms(int l0, int l1, int lane, double rm, float inv_m, float &vt, float &vg,
                                            float &vf);
template <typename MFn, typename RmFn>
__device__ __forceinline__ void efron_chunks(const int *cp, int nb, int gw, int W, int lane, MFn m_of, RmFn rm_of,
                                             long long *acc) {
    const int C = cp[nb], q = (C + W - 1) / W;
    int c = gw * q;
    const int c1 = min(C, c + q);
    if (c >= c1) return;
    int lo = 0, hi = nb;  // cp[lo] <= c < cp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cp[mid] <= c) lo = mid; else hi = mid;
    }
    int p = lo;
    while (c < c1) {
        const int base = cp[p], cend = min(c1, cp[p + 1]);
        if (cend > c) {
            const int b = nb - 1 - p, m = m_of(b);
            const double rm = rm_of(b);
            const float inv_m = __frcp_rn((float)m);
            long long it = 0, ig = 0, iff = 0;
            // one chunk = one unit of floating-point summation, converted to fixed point before it meets any other:
            // the per-bin sums do not depend on how the chunks are dealt out (grid size, sharding, code path)
            for (int l0 = (c - base) * EF_CHUNK; l0 < (cend - base) * EF_CHUNK; l0 += EF_CHUNK) {
                float vt = 0.f, vg = 0.f, vf = 0.f;
                efron_terms(l0, min(m, l0 + EF_CHUNK), lane, rm, inv_m, vt, vg, vf);
                vt = warp_sum(vt); vg = warp_sum(vg); vf = warp_sum(vf);
                it += __double2ll_rn((double)vt * EF_SCALE);
                ig += __double2ll_rn((double)vg * EF_SCALE);
                iff += __double2ll_rn((double)vf * EF_SCALE);
            }
            if (lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long *>(acc + b), (unsigned long long)it);
                atomicAdd(reinterpret_cast<unsigned long long *>(acc + nb + b), (unsigned long long)ig);
                atomicAdd(reinterpret_cast<unsigned long long *>(acc + 2 * nb + b), (unsigned long long)iff);
            }
            c = cend;
        }
        ++p;
    }
}

// ================================================================ K2: reduce partials (+ fused scan)
// grid (nb / 32, n_seg), 1024 threads = 32 bins x 32 groups of partials.
template <int MAXPER>
__global__ void __launch_bounds__(RED_THREADS)
cox_binned_reduce(const unsigned char *__restrict__ partial, const CtaRec *__restrict__ recs, int nctas, int nb,
                  long long *bins, float *__restrict__ bins_max, ScanBase sb, int fuse_scan) {
    __shared__ long long s_c[RED_NG][RED_BINS], s_e[RED_NG][RED_BINS];
    __shared__ unsigned s_m[RED_NG][RED_BINS];
    __shared__ double shd[33];
    __shared__ int shi[36];
    __shared__ int s_last;
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
        float mx = -INFINITY;
        unsigned fl = 0;
        for (int c = threadIdx.x; c < nctas; c += 32) {
            const CtaRec r = recs[(size_t)seg * nctas + c];
            se += r.sum_ev_eta; sw += r.sum_w; mx = fmaxf(mx, r.max_eta); fl |= r.flags;
        }
        se = warp_sum(se); sw = warp_sum(sw); mx = warp_max(mx); fl = warp_or(fl);
        if (threadIdx.x == 0) {
            bs[3 * (size_t)nb + 0] = __double2ll_rn(se * ETA_SCALE);
            bs[3 * (size_t)nb + 1] = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1 : 0;
            bs[3 * (size_t)nb + 2] = (long long)fmin(ceil(sw), 9.0e18);
            bs[3 * (size_t)nb + 3] = (fl & B200SURV_COXF_BAD_TIME) ? 1 : 0;
            bins_max[2 * seg + 0] = mx;
            bins_max[2 * seg + 1] = -1.f;  // reserved
        }
    }
    if (!fuse_scan) return;
    // ---- the last CTA of this segment to arrive scans the finished bins
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence();

        const unsigned prev = atomicAdd(sb.tickets_k2 + seg, 1u);  // zeroed by pass 1
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_segment<MAXPER>(bs, nb, scan_bufs(sb, seg, nb), shd, shi);
}

// stand-alone scan for the multi-GPU path (bins come out of an all-reduce): grid n_seg, 1024 threads
template <int MAXPER>
__global__ void __launch_bounds__(RED_THREADS)
cox_binned_scan(const long long *bins, int nb, ScanBase sb) {
    __shared__ double shd[33];
    __shared__ int shi[36];
    const int seg = blockIdx.x;
    scan_segment<MAXPER>(bins + (size_t)seg * (3 * (size_t)nb + 4), nb, scan_bufs(sb, seg, nb), shd, shi);
}

// ================================================================ K3: items + finish
// state per segment: header (64 B) | float2 (P, F)[nb]
__host__ __device__ inline size_t seg_state_stride(int nb) {
    return sizeof(b200surv_cox_header) + (size_t)nb * sizeof(float2);
}

// sum over l in [l0, l1), stride 32 per lane, of log(x), 1/x, (l/m)/x with x = 1 - l * rm
__device__ __forceinline__ void efron_terms(int l0, int l1, int lane, double rm, float inv_m, float &vt, float &vg,
                                            float &vf) {
    if ((double)(l1 - 1) * rm <= 0.5) {  // x >= 0.5: fp32 is accurate to ~1e-7 relative
        const float rmf = (float)rm;
        // sum of logs = log of the product: a call covers at most EF_CHUNK = 256 terms, 8 per lane, and 8 factors in
        // [0.5, 1] stay >= 2^-8 -- one MUFU.LG2 per lane instead of one per term
        float pr = 1.f;
#pragma unroll 2
        for (int l = l0 + lane; l < l1; l += 32) {
            const float x = fmaf(-(float)l, rmf, 1.f);
            const float rx = rcp_approx(x);
            pr *= x;
            vg += rx;
            vf = fmaf((float)l * inv_m, rx, vf);
        }
        vt += __logf(pr);
    } else {  // the events are a large part of the risk set: keep the difference in fp64
#pragma unroll 1
        for (int l = l0 + lane; l < l1; l += 32) {
            const double xd = 1.0 - (double)l * rm;
            const float x = (float)xd;
            const float rx = (float)(1.0 / xd);
            vt += __logf(x);
            vg += rx;
            vf = fmaf((float)l * inv_m, rx, vf);
        }
    }
}

// grid (gx, n_seg), 1024 threads, no shared-memory staging: warp gw owns bins gw, gw + W, ...
template <int MAXPER>
__global__ void __launch_bounds__(IT_THREADS, 1)
cox_binned_items_finish(const long long *__restrict__ bins, const float *__restrict__ bins_max, int nb, int ties,
                        int reduction, float shift, ScanBase sb, float *__restrict__ out_loss,
                        unsigned char *__restrict__ state) {
    __shared__ double shd[33];
    __shared__ int s_last;
    const int seg = blockIdx.y, t = threadIdx.x;
    const long long *bs = bins + (size_t)seg * (3 * (size_t)nb + 4);
    const ScanBufs o = scan_bufs(sb, seg, nb);
    const bool efron = ties == B200SURV_TIES_EFRON;

    if (efron) {
        __shared__ int s_cp[B200SURV_COX_MAX_BINS + 1];
        for (int i = t; i <= nb; i += IT_THREADS) s_cp[i] = o.cp[i];
        __syncthreads();
        // consecutive warp ids go to different CTAs (balance across SMs when the runs are short)
        const int gw = (t >> 5) * gridDim.x + blockIdx.x, W = gridDim.x * (IT_THREADS / 32);
        efron_chunks(
            s_cp, nb, gw, W, t & 31, [&](int b) { return (int)bs[2 * nb + b]; }, [&](int b) { return o.rm[b]; },
            reinterpret_cast<long long *>(o.tgf));
    }
    // ---- last CTA of the segment to arrive finishes
    __syncthreads();

    if (t == 0) { __threadfence(); s_last = (atomicAdd(o.ticket, 1u) == gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    unsigned char *seg_state = state + seg * seg_state_stride(nb);
    b200surv_cox_header *hdr = reinterpret_cast<b200surv_cox_header *>(seg_state);
    float2 *table = reinterpret_cast<float2 *>(seg_state + sizeof(b200surv_cox_header));
    const int per = nb / IT_THREADS > 0 ? nb / IT_THREADS : 1;
    const int lo_b = t * per;
    double tv[MAXPER], gv[MAXPER], fv[MAXPER];
    int mv[MAXPER];
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {  // tgf was written by other CTAs of this launch: read through L2
        const int b = lo_b + k;
        const bool in = (k < per) && (b < nb);
        const long long *acc = reinterpret_cast<const long long *>(o.tgf);
        mv[k] = in ? (int)bs[2 * nb + b] : 0;
        const bool live = mv[k] > 0;
        const long long at = live ? __ldcg(acc + b) : 0ll, ag = live ? __ldcg(acc + nb + b) : 0ll,
                        af = live ? __ldcg(acc + 2 * nb + b) : 0ll;
        tv[k] = gv[k] = fv[k] = 0.0;
        if (live) {  // the same expressions as the fused forward: bit-identical results on both paths
            const double D = o.D[b], invD = 1.0 / D, m = (double)mv[k];
            tv[k] = (double)at * EF_INV + m * log(D);
            gv[k] = efron ? (double)ag * EF_INV * invD : m * invD;
            fv[k] = (double)af * EF_INV * invD;
        }
    }
    double tsum = 0.0, gsum = 0.0;
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) { tsum += tv[k]; gsum += gv[k]; }
    __shared__ double shd2[33];
    double T;
    double run = block_exscan_and_sum(gsum, tsum, shd, shd2, &T);
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        const int b = lo_b + k;
        if (k < per && b < nb) { run += gv[k]; table[b] = make_float2((float)run, (float)fv[k]); }
    }
    if (t == 0) {
        const long long n_events = o.totals[0], n_times = o.totals[1];
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
        hdr->shift = shift; hdr->max_log_hz = mx; hdr->max_time = -1.f;
        hdr->nbins = nb; hdr->n_events = n_events; hdr->n_event_times = n_times; hdr->pll = pll;
        hdr->reserved = 0;
        out_loss[seg] = loss;
    }
}

// ================================================================ fused forward (one cohort, one launch)
// pass 1, the exact reduction of the CTA partials, the suffix scan, the Efron terms and the finish in ONE
// cooperative kernel (one CTA per SM, grid barriers between the phases): removes two launches and the two
// single-CTA tails of the K2/K3 pipeline.  Every CTA scans the nbins sums redundantly in shared memory.
struct FusedArgs {
    long long *bins;   // [3 nb + 4]
    float *bins_max;   // [2]
    double *tgf;       // [3][nb]
    int ties, reduction;
    float *out_loss;
    unsigned char *state;
};

// ---- multi-GPU exchange through peer memory (NVLink / NVSwitch), fused into the cooperative forward.
// Every rank owns one "peer buffer" that all ranks of the box have mapped (symmetric allocation):
//   [ flags: PEER_MAX_WORLD x 128 B, slot r is written by rank r ]
//   [ slot 0 | slot 1 ]   each: int64 bins[3 nb + 4], float max[2]     (slot = epoch & 1)
// Step k: a rank writes its own per-bin sums into its slot k&1, publishes flag[rank] = k in every peer's
// buffer (release, system scope), waits until its own flags show k for every peer, then sums the peers'
// slots bin slice by bin slice (each CTA pulls ~nb/148 bins from every peer: (world-1) * 24 nb bytes per GPU
// over NVLink in total).  Integer sums, so every rank obtains bit-identical totals.  Double buffering is
// enough: a rank rewrites slot k&1 in step k+2, after it passed the barrier of step k+1, which every peer
// only signals once its step-k kernel (the one reading this slot) has finished.
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_FLAG_STRIDE = 32;  // unsigned words (128 B)
constexpr size_t PEER_FLAGS_BYTES = (size_t)PEER_MAX_WORLD * PEER_FLAG_STRIDE * sizeof(unsigned);
constexpr long long PEER_SPIN_LIMIT_NS = 2000000000ll;  // 2 s: a missing peer must not hang the GPU

struct PeerArgs {
    int world, rank;
    unsigned epoch;
    int *status;  // local: set to 1 when the wait timed out
    long long *trace;  // optional (may be null): globaltimer stamps of the phases, CTA 0
    unsigned char *buf[PEER_MAX_WORLD];
};

__host__ __device__ inline size_t peer_slot_bytes(int nb) {
    return ((3 * (size_t)nb + 4) * sizeof(long long) + 2 * sizeof(float) + 255) / 256 * 256;
}

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_sys_s64(const long long *p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long global_timer_ns() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}

template <int MAXPER, bool PEER>
__global__ void __launch_bounds__(P1_THREADS, 1)
cox_binned_fwd_fused(const float *__restrict__ log_hz, const float *__restrict__ time,
                     const uint8_t *__restrict__ event, int64_t n, int nb, float shift, int vec_ok,
                     unsigned char *__restrict__ partial, CtaRec *__restrict__ recs, FusedArgs fa, int use_tma,
                     PeerArgs pa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double shd[33];
    __shared__ int shi[36];
    cg::grid_group grid = cg::this_grid();
    const int cta = blockIdx.x, nctas = gridDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (PEER && cta == 0 && t == 0) *pa.status = 0;  // set again (after a grid barrier) if a peer never arrives
#define PEER_TRACE(i)                                                                          \
    do {                                                                                       \
        if (PEER && pa.trace != nullptr && cta == 0 && t == 0) pa.trace[i] = global_timer_ns(); \
    } while (0)
    PEER_TRACE(0);
    if (use_tma == 2)
        pass1_body_ring(log_hz, time, event, n, nb, shift, partial, recs, cta, nctas, smem_raw);
    else if (use_tma == 1)
        pass1_body_tma(log_hz, time, event, n, nb, shift, partial, recs, cta, nctas, smem_raw);
    else
        pass1_body(log_hz, time, event, nullptr, n, nb, shift, vec_ok, partial, recs, 0, cta, nctas,
                   reinterpret_cast<unsigned *>(smem_raw));
    PEER_TRACE(1);
    grid.sync();
    PEER_TRACE(2);

    // ---- phase 2: exact reduction of the partials, bins [b0, b1) of this CTA, one warp per bin
    long long *bs = fa.bins;
    const size_t slot_off = PEER ? PEER_FLAGS_BYTES + (size_t)(pa.epoch & 1u) * peer_slot_bytes(nb) : 0;
    const int nbpc = (nb + nctas - 1) / nctas;
    {
        // with peers the rank-local sums go to this rank's slot of the peer buffer; fa.bins receives the totals
        long long *bs = PEER ? reinterpret_cast<long long *>(pa.buf[pa.rank] + slot_off) : fa.bins;
        float *bmax = PEER ? reinterpret_cast<float *>(bs + 3 * (size_t)nb + 4) : fa.bins_max;
        const int b0 = cta * nbpc, b1 = min(nb, b0 + nbpc);
        const bool wide = nctas >= 32 && nb >= 1024;  // (the shared histogram storage, free again, holds 20 KB)
        if (wide) {
            // lanes along the bins (contiguous 256-byte reads of every partial record), warps along the records,
            // then a cross-warp sum through shared memory
            unsigned long long *s_c = reinterpret_cast<unsigned long long *>(smem_raw);  // [32 warps][32 bins]
            unsigned long long *s_e = s_c + 1024;
            unsigned *s_mm = reinterpret_cast<unsigned *>(s_e + 1024);
            for (int g0 = b0; g0 < b1; g0 += 32) {
                const int b = g0 + lane;
                const bool bin_ok = b < b1;
                unsigned long long vc[RED_MAX_ITERS], ve[RED_MAX_ITERS];
                unsigned vm[RED_MAX_ITERS];
#pragma unroll
                for (int k = 0; k < RED_MAX_ITERS; ++k) {
                    const int c = warp + 32 * k;
                    const bool in = bin_ok && c < nctas;
                    const unsigned char *pc = partial + (size_t)(in ? c : 0) * PARTIAL_BYTES_PER_BIN * (size_t)nb;
                    vc[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + b) : 0ull;
                    ve[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + nb + b) : 0ull;
                    vm[k] = in ? __ldcg(reinterpret_cast<const unsigned *>(pc + 16 * (size_t)nb) + b) : 0u;
                }
                unsigned long long sc = 0, se = 0;
                unsigned m = 0;
#pragma unroll
                for (int k = 0; k < RED_MAX_ITERS; ++k) { sc += vc[k]; se += ve[k]; m += vm[k]; }
                if (g0 != b0) __syncthreads();  // the previous group's sums have been read
                s_c[t] = sc; s_e[t] = se; s_mm[t] = m;
                __syncthreads();
                if (warp < 3 && bin_ok) {  // warp 0: censored sums, warp 1: event sums, warp 2: event counts
                    long long tot = 0;
                    if (warp == 0) { for (int w = 0; w < 32; ++w) tot += (long long)s_c[w * 32 + lane]; }
                    else if (warp == 1) { for (int w = 0; w < 32; ++w) tot += (long long)s_e[w * 32 + lane]; }
                    else { for (int w = 0; w < 32; ++w) tot += (long long)s_mm[w * 32 + lane]; }
                    bs[(size_t)warp * nb + b] = tot;
                    fa.tgf[(size_t)warp * nb + b] = 0.0;
                }
            }
        }
        for (int b = b0 + warp; b < b1 && !wide; b += P1_THREADS / 32) {
            unsigned long long vc[RED_MAX_ITERS], ve[RED_MAX_ITERS];
            unsigned vm[RED_MAX_ITERS];
#pragma unroll
            for (int k = 0; k < RED_MAX_ITERS; ++k) {  // all loads in flight together (nctas <= 160)
                const int c = lane + 32 * k;
                const bool in = c < nctas;
                const unsigned char *pc = partial + (size_t)(in ? c : 0) * PARTIAL_BYTES_PER_BIN * (size_t)nb;
                vc[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + b) : 0ull;
                ve[k] = in ? __ldcg(reinterpret_cast<const unsigned long long *>(pc) + nb + b) : 0ull;
                vm[k] = in ? __ldcg(reinterpret_cast<const unsigned *>(pc + 16 * (size_t)nb) + b) : 0u;
            }
            unsigned long long sc = 0, se = 0;
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < RED_MAX_ITERS; ++k) { sc += vc[k]; se += ve[k]; m += vm[k]; }
            const long long tc = warp_sum((long long)sc), te = warp_sum((long long)se), tm = warp_sum((long long)m);
            if (lane == 0) {
                bs[b] = tc; bs[nb + b] = te; bs[2 * nb + b] = tm;
                fa.tgf[b] = 0.0; fa.tgf[nb + b] = 0.0; fa.tgf[2 * nb + b] = 0.0;
            }
        }
        if (cta == 0 && warp == P1_THREADS / 32 - 1) {
            double se = 0.0, sw = 0.0;
            float mx = -INFINITY;
            unsigned fl = 0;
            for (int c = lane; c < nctas; c += 32) {
                const CtaRec *rp = recs + c;
                se += __ldcg(&rp->sum_ev_eta); sw += __ldcg(&rp->sum_w); mx = fmaxf(mx, __ldcg(&rp->max_eta));
                fl |= __ldcg(&rp->flags);
            }
            se = warp_sum(se); sw = warp_sum(sw); mx = warp_max(mx); fl = warp_or(fl);
            if (lane == 0) {
                bs[3 * (size_t)nb + 0] = __double2ll_rn(se * ETA_SCALE);
                bs[3 * (size_t)nb + 1] = (fl & B200SURV_COXF_NOT_BINNABLE) ? 1 : 0;
                bs[3 * (size_t)nb + 2] = (long long)fmin(ceil(sw), 9.0e18);
                bs[3 * (size_t)nb + 3] = (fl & B200SURV_COXF_BAD_TIME) ? 1 : 0;
                bmax[0] = mx;
                bmax[1] = -1.f;
            }
        }
    }
    PEER_TRACE(3);
    grid.sync();
    PEER_TRACE(4);

    if constexpr (PEER) {
        // ---- phase 2x: publish, wait for every peer, pull and add their sums for this CTA's bin slice
        __shared__ int s_timeout;
        if (t == 0) s_timeout = 0;
        if (cta == 0 && t < pa.world) {
            __threadfence_system();  // the slot written by all CTAs (ordered by the grid barrier) before the flag
            st_release_sys_u32(reinterpret_cast<unsigned *>(pa.buf[t]) + pa.rank * PEER_FLAG_STRIDE, pa.epoch);
        }
        __syncthreads();
        PEER_TRACE(5);
        if (t < pa.world) {
            const unsigned *f = reinterpret_cast<const unsigned *>(pa.buf[pa.rank]) + t * PEER_FLAG_STRIDE;
            const long long t0 = global_timer_ns();
            while ((int)(ld_acquire_sys_u32(f) - pa.epoch) < 0) {
                if (global_timer_ns() - t0 > PEER_SPIN_LIMIT_NS) { s_timeout = 1; break; }
                __nanosleep(64);
            }
        }
        __syncthreads();
        PEER_TRACE(6);
        if (s_timeout) {  // a peer never arrived: flag it and leave an empty cohort behind (no stale sums)
            if (t == 0) atomicExch(pa.status, 1);
            const int b0 = cta * nbpc, b1 = min(nb, b0 + nbpc);
            for (int b = b0 + t; b < b1; b += P1_THREADS) { bs[b] = 0; bs[nb + b] = 0; bs[2 * (size_t)nb + b] = 0; }
            if (cta == nctas - 1 && t < 4) bs[3 * (size_t)nb + t] = 0;
            if (cta == nctas - 1 && t == 4) { fa.bins_max[0] = 0.f; fa.bins_max[1] = -1.f; }
        } else {
            const int b0 = cta * nbpc, b1 = min(nb, b0 + nbpc);
            const int cnt = max(b1 - b0, 0);
            for (int i = t; i < 3 * cnt; i += P1_THREADS) {
                const size_t idx = (size_t)(i / cnt) * nb + b0 + (i % cnt);
                long long v[PEER_MAX_WORLD];
#pragma unroll
                for (int p = 0; p < PEER_MAX_WORLD; ++p)  // all peers' loads in flight together
                    v[p] = p < pa.world ? ld_relaxed_sys_s64(reinterpret_cast<const long long *>(pa.buf[p] + slot_off) + idx) : 0ll;
                long long s = 0;
#pragma unroll
                for (int p = 0; p < PEER_MAX_WORLD; ++p) s += v[p];
                bs[idx] = s;
            }
            if (cta == nctas - 1 && t >= P1_THREADS - 5) {  // the four scalar words and the maximum
                const int k = t - (P1_THREADS - 5);
                if (k < 4) {
                    long long s = 0;
                    for (int p = 0; p < pa.world; ++p)
                        s += ld_relaxed_sys_s64(reinterpret_cast<const long long *>(pa.buf[p] + slot_off) + 3 * (size_t)nb + k);
                    bs[3 * (size_t)nb + k] = s;
                } else {
                    float mx = -INFINITY;
                    for (int p = 0; p < pa.world; ++p)
                        mx = fmaxf(mx, ld_relaxed_sys_f32(reinterpret_cast<const float *>(
                                           reinterpret_cast<const long long *>(pa.buf[p] + slot_off) + 3 * (size_t)nb + 4)));
                    fa.bins_max[0] = mx;
                    fa.bins_max[1] = -1.f;
                }
            }
        }
        grid.sync();
        PEER_TRACE(7);
    }

    // ---- phase 3: redundant suffix scan into shared memory (reuses the histogram storage: 20 B/bin)
    double *sD = reinterpret_cast<double *>(smem_raw);
    double *s_rm = sD + nb;
    int *s_m = reinterpret_cast<int *>(s_rm + nb);
    // s_cp[p], p = nb-1-b (scan order): exclusive prefix of the Efron chunk counts ceil(m/EF_CHUNK); [nb] = total
    // (the fused launch allocates 24 B/bin + 16)
    int *s_cp = s_m + nb;
    const bool efron = fa.ties == B200SURV_TIES_EFRON;
    const bool solo = nctas > 1;
    const bool skip_efron = !efron || (solo && cta == 0);
    const int per = nb / P1_THREADS > 0 ? nb / P1_THREADS : 1;
    const int hi_b = nb - t * per;
    int n_events, n_times;
    {
        long long rc[MAXPER], re[MAXPER], rmv[MAXPER];
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = hi_b - 1 - k;
            const bool in = (k < per) && (b >= 0);
            rc[k] = in ? __ldcg(bs + b) : 0ll;
            re[k] = in ? __ldcg(bs + nb + b) : 0ll;
            rmv[k] = in ? __ldcg(bs + 2 * nb + b) : 0ll;
        }
        double sv[MAXPER];
        double loc = 0.0;
        int locm = 0, net = 0, locc = 0;
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            sv[k] = ((double)(unsigned long long)rc[k] + (double)(unsigned long long)re[k]) * FIX_INV;
            loc += sv[k];
            locm += (int)rmv[k];
            net += rmv[k] > 0 ? 1 : 0;
            locc += ((int)rmv[k] + EF_CHUNK - 1) / EF_CHUNK;
        }
        const BinScan sc = block_bin_scan(loc, locc, locm, net, shd, shi);
        n_events = sc.n_events; n_times = sc.n_times;
        double run = sc.run;
        int crun = sc.crun;
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = hi_b - 1 - k;
            if (k < per && b >= 0) {
                run += sv[k];
                const int m = (int)rmv[k];
                sD[b] = run;
                s_m[b] = m;
                if (!skip_efron) {  // (CTA 0 of a multi-CTA grid takes no Efron work: it is on the critical path)
                    s_rm[b] = m > 0 ? ((double)(unsigned long long)re[k] * FIX_INV) / (run * (double)m) : 0.0;
                    s_cp[t * per + k] = crun;
                    crun += (m + EF_CHUNK - 1) / EF_CHUNK;
                }
            }
        }
        if (t == 0) s_cp[nb] = sc.n_chunks;
    }
    __syncthreads();
    // CTA 0 finishes alone after the last grid barrier; while the other CTAs work through the Efron terms it prepares
    // what does not depend on them (m log D and 1/D of its bins, the scalar words)
    const int lo_b = t * per;
    double pre_ml[MAXPER], pre_inv[MAXPER];
    long long sc_eta = 0, sc_nb = 0, sc_sw = 0, sc_bt = 0;
    float sc_mx = 0.f;
    if (cta == 0) {
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int b = lo_b + k;
            const bool in = (k < per) && (b < nb) && s_m[b] > 0;
            const double D = in ? sD[b] : 1.0;
            pre_inv[k] = 1.0 / D;
            pre_ml[k] = in ? (double)s_m[b] * log(D) : 0.0;
        }
        if (t == 0) {
            sc_eta = __ldcg(bs + 3 * (size_t)nb); sc_nb = __ldcg(bs + 3 * (size_t)nb + 1);
            sc_sw = __ldcg(bs + 3 * (size_t)nb + 2); sc_bt = __ldcg(bs + 3 * (size_t)nb + 3);
            sc_mx = __ldcg(fa.bins_max);
        }
    }
    if (PEER && pa.trace != nullptr && cta == 1 && t == 0) pa.trace[13] = global_timer_ns();
    if (efron && !(solo && cta == 0)) {
        const int wc = solo ? nctas - 1 : 1, ci = solo ? cta - 1 : 0;
        efron_chunks(
            s_cp, nb, warp * wc + ci, wc * (P1_THREADS / 32), lane, [&](int b) { return s_m[b]; },
            [&](int b) { return s_rm[b]; }, reinterpret_cast<long long *>(fa.tgf));
    }
    if (PEER && pa.trace != nullptr && cta == 1 && t == 0) pa.trace[14] = global_timer_ns();
    if (cta != 0) { grid.sync(); return; }

    // ---- phase 4 (CTA 0): P = prefix(G), loss, header, (P,F) table
    // The finish is ~500 instructions that run once per launch on one SM: executed cold it is bound by instruction
    // fetch (every 128-byte line a serial L2 round trip; measured 9 us for ~2 us of work).  So CTA 0, which has
    // nothing else to do while the other CTAs work through the Efron terms, runs it twice: a rehearsal on whatever the
    // accumulators hold (stores suppressed) that pulls the code into the SM's instruction cache, then, after the
    // grid barrier, the real pass.
    b200surv_cox_header *hdr = reinterpret_cast<b200surv_cox_header *>(fa.state);
    float2 *table = reinterpret_cast<float2 *>(fa.state + sizeof(b200surv_cox_header));
    const int npass = solo ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
    const bool live = pass == npass - 1;
    if (live) {
        PEER_TRACE(8);
        grid.sync();
        PEER_TRACE(9);
    }
    double tv[MAXPER], gv[MAXPER], fv[MAXPER];
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        const int b = lo_b + k;
        const bool in = (k < per) && (b < nb) && s_m[b] > 0;
        const long long *acc = reinterpret_cast<const long long *>(fa.tgf);
        const long long at = in ? __ldcg(acc + b) : 0ll, ag = in ? __ldcg(acc + nb + b) : 0ll,
                        af = in ? __ldcg(acc + 2 * nb + b) : 0ll;
        tv[k] = gv[k] = fv[k] = 0.0;
        if (in) {
            const double invD = pre_inv[k], m = (double)s_m[b];
            tv[k] = (double)at * EF_INV + pre_ml[k];
            gv[k] = efron ? (double)ag * EF_INV * invD : m * invD;
            fv[k] = (double)af * EF_INV * invD;
        }
    }
    double tsum = 0.0, gsum = 0.0;
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) { tsum += tv[k]; gsum += gv[k]; }
    if (PEER && pa.trace != nullptr && t == 0) pa.trace[11] = global_timer_ns() + (tsum > 1e300 ? 1 : 0);
    __shared__ double shd2[33];
    double T;
    double run = block_exscan_and_sum(gsum, tsum, shd, shd2, &T);
    if (PEER && pa.trace != nullptr && t == 0) pa.trace[12] = global_timer_ns() + (run > 1e300 ? 1 : 0);
    // the rehearsal stores into shared scratch (the D array is no longer needed; shd2 after its last read)
    float2 *table_w = live ? table : reinterpret_cast<float2 *>(sD);
    b200surv_cox_header *hdr_w = live ? hdr : reinterpret_cast<b200surv_cox_header *>(shd2);
    float *loss_w = live ? fa.out_loss : reinterpret_cast<float *>(shd2 + 16);
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        const int b = lo_b + k;
        if (k < per && b < nb) { run += gv[k]; table_w[b] = make_float2((float)run, (float)fv[k]); }
    }
    if (t == 0) {
        const double sum_ev_eta = (double)sc_eta * ETA_INV;
        const double pll = sum_ev_eta - (T + (double)n_events * (double)shift);
        double norm = 1.0;
        if (fa.reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)n_events;
        else if (fa.reduction == B200SURV_REDUCE_MEAN_TERMS) norm = efron ? (double)n_times : (double)n_events;
        unsigned flags = 0;
        if (sc_nb != 0) flags |= B200SURV_COXF_NOT_BINNABLE;
        if (sc_bt != 0) flags |= B200SURV_COXF_BAD_TIME;
        const float mx = sc_mx;
        const double sumw = (double)sc_sw;
        if (!(mx - shift <= SHIFT_HI) || !(mx - shift >= SHIFT_LO) || sumw >= SUMW_LIMIT) flags |= B200SURV_COXF_EXP_RANGE;
        if (PEER && __ldcg(pa.status) != 0) flags |= B200SURV_COXF_PEER_TIMEOUT;
        float loss = 0.f, scale = 0.f;
        if (n_events > 0) { loss = (float)(-pll / norm); scale = (float)(-1.0 / norm); }
        if (flags) { loss = __int_as_float(0x7fc00000); scale = loss; }
        hdr_w->flags = flags; hdr_w->mode = B200SURV_COX_BINNED; hdr_w->loss = loss; hdr_w->scale = scale;
        hdr_w->shift = shift; hdr_w->max_log_hz = mx; hdr_w->max_time = -1.f;
        hdr_w->nbins = nb; hdr_w->n_events = n_events; hdr_w->n_event_times = n_times; hdr_w->pll = pll;
        hdr_w->reserved = 0;
        loss_w[0] = loss;
        PEER_TRACE(10);
    }
    __syncthreads();  // rehearsal scratch (shd2) is reused by the live pass
    }
#undef PEER_TRACE
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
    size_t off_partial, off_recs, off_tickets, off_bins, off_bins_max, off_D, off_rm, off_tgf, off_totals, off_big,
        total;
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
    // the tickets come first: the caller-visible contract is that a fresh workspace starts zero-filled
    // in its first 256-byte-aligned block of 2 * n_seg words (see cox_binned_workspace_init_bytes)
    L.off_tickets = take((size_t)n_seg * 2 * sizeof(unsigned));
    L.off_partial = take((size_t)n_seg * L.nctas * PARTIAL_BYTES_PER_BIN * nb);
    L.off_recs = take((size_t)n_seg * L.nctas * sizeof(CtaRec));
    L.off_bins = take((size_t)n_seg * (3 * (size_t)nb + 4) * sizeof(long long));
    L.off_bins_max = take((size_t)n_seg * 2 * sizeof(float));
    L.off_D = take((size_t)n_seg * nb * sizeof(double));
    L.off_rm = take((size_t)n_seg * nb * sizeof(double));
    L.off_tgf = take((size_t)n_seg * 3 * nb * sizeof(double));
    L.off_totals = take((size_t)n_seg * 2 * sizeof(long long));
    L.off_big = take((size_t)n_seg * (nb + 1) * sizeof(int));
    L.total = o;
    return L;
}

ScanBase make_scan_base(const BinnedLayout &L, unsigned char *w8, int64_t n_seg) {
    ScanBase s;
    s.D = reinterpret_cast<double *>(w8 + L.off_D);
    s.rm = reinterpret_cast<double *>(w8 + L.off_rm);
    s.tgf = reinterpret_cast<double *>(w8 + L.off_tgf);
    s.totals = reinterpret_cast<long long *>(w8 + L.off_totals);
    s.big = reinterpret_cast<int *>(w8 + L.off_big);
    s.tickets_k2 = reinterpret_cast<unsigned *>(w8 + L.off_tickets);
    s.tickets_k3 = s.tickets_k2 + n_seg;
    return s;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool coop_supported() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, v = 0;
        cached = (cudaGetDevice(&dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v) ? 1 : 0;
    }
    return cached == 1;
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
                            const BinnedLayout &L, unsigned char *w8, int fuse_scan, cudaStream_t st) {
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
    const ScanBase sb = make_scan_base(L, w8, n_seg);
    cox_binned_pass1<<<dim3(L.nctas, (unsigned)n_seg), P1_THREADS, smem, st>>>(
        log_hz, time, event, seg_off, n, nb, shift, vec_ok, partial, recs, sb.tickets_k2);
    const dim3 grid(nb / RED_BINS, (unsigned)n_seg);
    if (nb <= 4 * RED_THREADS)
        cox_binned_reduce<4><<<grid, RED_THREADS, 0, st>>>(partial, recs, L.nctas, nb, bins, bins_max, sb, fuse_scan);
    else
        cox_binned_reduce<8><<<grid, RED_THREADS, 0, st>>>(partial, recs, L.nctas, nb, bins, bins_max, sb, fuse_scan);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t launch_items_finish(const long long *bins, const float *bins_max, int64_t n_seg, int ties, int reduction,
                            int nb, float shift, float *out_loss, void *state, const BinnedLayout &L,
                            unsigned char *w8, int run_scan, cudaStream_t st) {
    const ScanBase sb = make_scan_base(L, w8, n_seg);
    if (run_scan) {
        if (nb <= 4 * RED_THREADS) cox_binned_scan<4><<<(unsigned)n_seg, RED_THREADS, 0, st>>>(bins, nb, sb);
        else cox_binned_scan<8><<<(unsigned)n_seg, RED_THREADS, 0, st>>>(bins, nb, sb);
    }
    int gx = 1;
    if (ties == B200SURV_TIES_EFRON) {
        gx = n_seg == 1 ? num_sms() : (int)((num_sms() + n_seg - 1) / n_seg);
        if (gx < 1) gx = 1;
    }
    const dim3 grid(gx, (unsigned)n_seg);
    if (nb <= 4 * IT_THREADS)
        cox_binned_items_finish<4><<<grid, IT_THREADS, 0, st>>>(bins, bins_max, nb, ties, reduction, shift, sb,
                                                                out_loss, static_cast<unsigned char *>(state));
    else
        cox_binned_items_finish<8><<<grid, IT_THREADS, 0, st>>>(bins, bins_max, nb, ties, reduction, shift, sb,
                                                                out_loss, static_cast<unsigned char *>(state));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t launch_fused(const float *log_hz, const float *time, const uint8_t *event, int64_t n, int ties, int reduction,
                     int nb, float shift, float *out_loss, void *state, const BinnedLayout &L, unsigned char *w8,
                     const PeerArgs *peer, cudaStream_t st);

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
                               reinterpret_cast<long long *>(bins_sum), bins_max, L, static_cast<unsigned char *>(ws),
                               /*fuse_scan=*/0, st);
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
                               shift, out_loss, state, L, static_cast<unsigned char *>(ws), /*run_scan=*/1, st);
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
    if (n_seg == 1 && seg_off == nullptr && coop_supported())
        return launch_fused(log_hz, time, event, n, ties, reduction, nb, shift, out_loss, state, L, w8, nullptr, st);
    rc = launch_pass1_reduce(log_hz, time, event, seg_off, n, n_seg, nb, shift, bins, bins_max, L, w8,
                             /*fuse_scan=*/1, st);
    if (rc) return rc;
    return launch_items_finish(bins, bins_max, n_seg, ties, reduction, nb, shift, out_loss, state, L, w8,
                               /*run_scan=*/0, st);
}

size_t cox_binned_peer_buffer_bytes(int nb) { return PEER_FLAGS_BYTES + 2 * peer_slot_bytes(nb); }
size_t cox_binned_peer_trace_offset(int64_t n, int nb) { return binned_layout(n, 1, nb).off_D; }

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
    pa.status = reinterpret_cast<int *>(w8 + L.off_totals);  // unused by the fused path otherwise
    static const bool trace_on = getenv("B200SURV_PEER_TRACE") != nullptr;
    pa.trace = trace_on ? reinterpret_cast<long long *>(w8 + L.off_D) : nullptr;  // 11 stamps, see cox_binned_peer_trace_offset
    for (int p = 0; p < world; ++p) {
        B200_REQUIRE(peer_bufs[p] != nullptr && (reinterpret_cast<uintptr_t>(peer_bufs[p]) & 255) == 0, "peer buffer alignment (256)");
        pa.buf[p] = static_cast<unsigned char *>(peer_bufs[p]);
    }
    return launch_fused(log_hz, time, event, n, ties, reduction, nb, shift, out_loss, state, L, w8, &pa, st);
}

namespace {
int32_t launch_fused(const float *log_hz, const float *time, const uint8_t *event, int64_t n, int ties, int reduction,
                     int nb, float shift, float *out_loss, void *state, const BinnedLayout &L, unsigned char *w8,
                     const PeerArgs *peer, cudaStream_t st) {
    long long *bins = reinterpret_cast<long long *>(w8 + L.off_bins);
    float *bins_max = reinterpret_cast<float *>(w8 + L.off_bins_max);
    {
        // one cooperative launch: pass 1 + reduce (+ peer exchange) + scan + Efron terms + finish
        int vec_ok = aligned16(log_hz) && aligned16(time) && ((reinterpret_cast<uintptr_t>(event) & 3) == 0);
        // The TMA-staged pass 1 (pass1_body_tma) is opt-in: measured 78.9 us vs 72.8 us for the register-staged
        // loop at 16.7M rows (profiles/r1_v10_tma_ab.txt) -- with 31 consumer warps the ring hides the latency
        // but the CTA loses a warp and pays an mbarrier round trip per 4 rows/thread.
        // Pass-1 staging, B200SURV_P1 = ring (default) | reg | tma.  Measured at 16.7M rows, forward only:
        // per-thread cp.async ring 64.4 us, register-staged loads 70.3 us, TMA bulk ring + producer warp 78.9 us
        // (profiles/r1_v12_p1_staging_ab.txt, r1_v10_tma_ab.txt).
        static const int p1_mode = [] {
            const char *e = getenv("B200SURV_P1");
            if (e && !strcmp(e, "tma")) return 1;
            if (e && !strcmp(e, "reg")) return 0;
            return 2;
        }();
        int use_tma = (vec_ok && nb <= 4096 && n >= 4 * (int64_t)TMA_TILE) ? p1_mode : 0;
        const size_t smem = use_tma == 2 ? ring_smem_bytes(nb) : use_tma == 1 ? tma_smem_bytes(nb) : (size_t)nb * 24 + 16;
        static bool attr_done = false;
        if (!attr_done) {
            size_t mx = (size_t)B200SURV_COX_MAX_BINS * 24 + 16;
            if (tma_smem_bytes(4096) > mx) mx = tma_smem_bytes(4096);
            if (ring_smem_bytes(4096) > mx) mx = ring_smem_bytes(4096);
            B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
            B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
            B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
            B200_CHECK_CUDA(cudaFuncSetAttribute(cox_binned_fwd_fused<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
            attr_done = true;
        }
        unsigned char *partial = w8 + L.off_partial;
        CtaRec *recs = reinterpret_cast<CtaRec *>(w8 + L.off_recs);
        FusedArgs fa;
        fa.bins = bins; fa.bins_max = bins_max; fa.tgf = reinterpret_cast<double *>(w8 + L.off_tgf);
        fa.ties = ties; fa.reduction = reduction; fa.out_loss = out_loss; fa.state = static_cast<unsigned char *>(state);
        int nb_i = nb;
        PeerArgs pa;
        if (peer) pa = *peer; else memset(&pa, 0, sizeof(pa));
        void *args[] = {(void *)&log_hz, (void *)&time, (void *)&event, (void *)&n, (void *)&nb_i, (void *)&shift,
                        (void *)&vec_ok, (void *)&partial, (void *)&recs, (void *)&fa, (void *)&use_tma, (void *)&pa};
        const bool small = nb <= 4 * P1_THREADS;
        const void *fn = peer ? (small ? (const void *)cox_binned_fwd_fused<4, true> : (const void *)cox_binned_fwd_fused<8, true>)
                              : (small ? (const void *)cox_binned_fwd_fused<4, false> : (const void *)cox_binned_fwd_fused<8, false>);
        B200_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(L.nctas), dim3(P1_THREADS), args, smem, st));
        return B200SURV_OK;
    }
}
}  // namespace

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
