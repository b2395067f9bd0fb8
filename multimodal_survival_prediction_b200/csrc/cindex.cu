// Harrell's concordance index: int64 pair counts on the GPU (tiled O(n^2) pair counting).
//
// Replaces torchsurv.metrics.cindex.ConcordanceIndex()(estimate, event, time) as called by the
// reference at scripts/training/partial_modality_training.py:290-294 and simple_fusion.py:330-331;
// pair rule and the six counters: include/b200surv.h, oracle/cindex_oracle.c.
//
// algo 0: literal all-pairs tiles, no preprocessing (kept as an independent cross-check).
// algo 1: sort rows by (time, events first).  Then the comparable set of an event row is a SUFFIX of
//   the sorted order: same-time censored rows [s, ge) followed by all later rows [ge, n).  Event rows
//   selected by [row_begin,row_end) are compacted, their tie thresholds lo/hi (tie <=> lo <= e_j <= hi,
//   found by bisection on the exact fp32 predicate) precomputed, and a tiled kernel counts
//   #(e_j < lo) and #(e_j <= hi) over upper-triangular tiles only.  Column tiles are staged in shared
//   memory and broadcast; each thread keeps 8 rows in registers; per-thread 32-bit counters are
//   reduced with warp shuffles into int64 atomics once per CTA.
// algo 2: algo 1's preprocessing, then ranks instead of pairs.  Every full column tile is ALSO kept sorted by estimate
//   (k_ci_tile_sort: one bitonic sort of 1024 floats per tile, NaN last), and the thresholds of every row tile are also kept
//   sorted (k_ci_row_sort).  Where a whole column tile is strictly later than every row of the row tile -- all tiles but the
//   handful around the row tile's own time span -- the sums over the row tile of #(e_j < lo) and #(e_j <= hi) are bisections
//   in shared memory instead of 2 x 1024 compares per row, taken from the column side (1024 sorted columns against the row
//   tile's 2048 sorted lo / hi thresholds); the searches of neighbouring lanes walk the same path because their values are
//   neighbours in sorted order (the sums over a row tile do not care which row a threshold belongs to).
//   In the tiles around the diagonal the rows keep their time order: the columns before a row's strictly-later range are
//   visited pair by pair and taken off the whole-tile ranks.  The same six integers (ranks in a sorted tile ARE the pair
//   counts); the partial last column tile is counted pair by pair as in algo 1.
// The radix sort and the prefix sum of the preprocessing are hand-written (sortscan.cuh: stable LSD radix sort,
// single-pass scan with decoupled look-back).
#include <climits>

#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

__device__ __forceinline__ bool is_tie(float a, float b, float tol) { return fabsf(__fsub_rn(a, b)) <= tol; }

// ------------------------------------------------------------------ algo 0: direct
constexpr int A0_THREADS = 256;
constexpr int A0_TILE = 1024;

__global__ void __launch_bounds__(A0_THREADS)
cindex_direct(const float *__restrict__ est, const float *__restrict__ time,
              const uint8_t *__restrict__ event, int64_t n, int64_t row_begin, int64_t row_end,
              float tol, unsigned long long *__restrict__ out) {
    __shared__ float s_e[A0_TILE], s_t[A0_TILE];
    __shared__ uint8_t s_v[A0_TILE];
    __shared__ long long red[32];
    const int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < row_end && event[i] != 0;
    const float ei = live ? est[i] : 0.f, ti = live ? time[i] : 0.f;
    long long c[6] = {0, 0, 0, 0, 0, 0};
    const int64_t c_begin = (int64_t)blockIdx.y * A0_TILE * 16, c_end = min(n, c_begin + (int64_t)A0_TILE * 16);
    for (int64_t c0 = c_begin; c0 < c_end; c0 += A0_TILE) {
        __syncthreads();
        for (int k = threadIdx.x; k < A0_TILE; k += blockDim.x) {
            const int64_t j = c0 + k;
            s_e[k] = j < n ? est[j] : 0.f;
            s_t[k] = j < n ? time[j] : -INFINITY;  // never comparable: not > ti, not == ti (ti >= 0)
            s_v[k] = j < n ? event[j] : 1;
        }
        __syncthreads();
        if (live) {
            const int lim = (int)min((int64_t)A0_TILE, n - c0);
            for (int k = 0; k < lim; ++k) {
                const float tj = s_t[k], ej = s_e[k];
                const bool strict = tj > ti;
                const bool same = (tj == ti) && !s_v[k];
                if (strict || same) {
                    const bool tie = is_tie(ei, ej, tol);
                    const bool conc = !tie && (ej < ei);
                    const int base = strict ? 0 : 3;
                    c[base + (tie ? 2 : (conc ? 0 : 1))] += 1;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const long long v = block_reduce<long long>(c[q], 0ll, OpAddLL(), red);
        if (threadIdx.x == 0 && v) atomicAdd(out + q, (unsigned long long)v);
    }
}

// ------------------------------------------------------------------ algo 1: sorted suffix counting
constexpr int CT_THREADS = 256;
constexpr int CT_R = 8;                       // rows per thread
constexpr int CT_ROWS = CT_THREADS * CT_R;    // 2048 rows per CTA
constexpr int CT_TILE = 1024;                 // columns per shared-memory tile

struct Acc1 {
    unsigned long long conc_s, le_s, conc_t, le_t, tot_s, tot_t;
    unsigned long long n_rows, work;  // work: the counter the count kernel's CTAs pull their items from
    unsigned long long work2;         // algo 2: the second pass over the items
};

__device__ __forceinline__ uint32_t time_key(float t, bool ev) {
    return (__float_as_uint(t + 0.f) << 1) | (ev ? 0u : 1u);
}
__device__ __forceinline__ uint32_t f2o(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(256)
k_ci_keys(const float *__restrict__ time, const uint8_t *__restrict__ event, int64_t n,
          uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, Acc1 *acc) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *acc = Acc1{0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = time_key(time[i], event[i] != 0);
        vals[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256)
k_ci_flag(const float *__restrict__ est, const uint32_t *__restrict__ keys_s,
          const uint32_t *__restrict__ idx_s, int64_t n, int64_t row_begin, int64_t row_end,
          float *__restrict__ est_s, int *__restrict__ isrow) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t i = idx_s[p];
        est_s[p] = est[i];
        isrow[p] = (!(keys_s[p] & 1u) && (int64_t)i >= row_begin && (int64_t)i < row_end) ? 1 : 0;
    }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t *a, int n, uint32_t v) {  // first a[q] >= v
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1); if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// per selected event row: comparable range [s, ge | ge, n), tie thresholds, row totals
__global__ void __launch_bounds__(256)
k_ci_rows(const uint32_t *__restrict__ keys_s, const float *__restrict__ est_s,
          const int *__restrict__ isrow, const int *__restrict__ rank, int64_t n, float tol,
          float *__restrict__ r_lo, float *__restrict__ r_hi, int *__restrict__ r_s,
          int *__restrict__ r_ge, int shard, int n_shards, Acc1 *acc) {
    __shared__ long long red[32];
    long long tot_s = 0, tot_t = 0, nrows = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        if (!isrow[p]) continue;
        const int k = rank[p];
        const uint32_t key = keys_s[p], grp = key >> 1;
        const int s = lower_bound_u32(keys_s, (int)n, (grp << 1) | 1u);
        // grp + 1 cannot overflow 31 bits for finite non-negative times
        const int ge = lower_bound_u32(keys_s, (int)n, (grp + 1u) << 1);
        const float e = est_s[p];
        float lo = e, hi = e;
        if (!is_tie(e, e, tol)) {        // +-inf or NaN: nothing ties with it, not even itself
            lo = e;                      // conc <=> e_j < e
            hi = (e == e) ? ((e > 0.f) ? 3.402823466e+38f : __int_as_float(0x7fc00000)) : e;  // e_j <= hi <=> e_j < e
        } else {
            uint32_t a = f2o(-INFINITY), b = f2o(e);  // smallest o in [a,b] with tie
            while (a < b) { const uint32_t mid = a + ((b - a) >> 1); if (is_tie(e, o2f(mid), tol)) b = mid; else a = mid + 1; }
            lo = o2f(a);
            a = f2o(e); b = f2o(INFINITY);            // largest o in [a,b] with tie
            while (a < b) { const uint32_t mid = a + ((b - a + 1) >> 1); if (is_tie(e, o2f(mid), tol)) a = mid; else b = mid - 1; }
            hi = o2f(a);
        }
        r_lo[k] = lo; r_hi[k] = hi; r_s[k] = s; r_ge[k] = ge;
        if ((k / CT_ROWS) % n_shards == shard) {  // the row tiles are dealt out round-robin to the shards
            tot_s += (long long)n - ge;
            tot_t += (long long)ge - s;
        }
        nrows += 1;
    }
    tot_s = block_reduce<long long>(tot_s, 0ll, OpAddLL(), red);
    tot_t = block_reduce<long long>(tot_t, 0ll, OpAddLL(), red);
    nrows = block_reduce<long long>(nrows, 0ll, OpAddLL(), red);
    if (threadIdx.x == 0) {
        if (tot_s) atomicAdd(&acc->tot_s, (unsigned long long)tot_s);
        if (tot_t) atomicAdd(&acc->tot_t, (unsigned long long)tot_t);
        if (nrows) atomicAdd(&acc->n_rows, (unsigned long long)nrows);
    }
}

// c += (a < b) / (a <= b) (ordered compares: false on NaN, like the C operators)
__device__ __forceinline__ void inc_lt(unsigned &c, float a, float b) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "f"(a), "f"(b));
}
__device__ __forceinline__ void inc_le(unsigned &c, float a, float b) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "f"(a), "f"(b));
}

// Persistent CTAs pull work items -- (row tile of this shard, column tile) pairs, 2048 rows x 1024 columns -- from an
// atomic counter: the number of event rows is only known on the device, and with a fixed grid of multi-tile CTAs the
// last, partly filled wave cost a 1/8 shard 1 ms of its 4 ms.
__global__ void __launch_bounds__(CT_THREADS)
k_ci_count(const float *__restrict__ est_s, int64_t n, const float *__restrict__ r_lo,
           const float *__restrict__ r_hi, const int *__restrict__ r_s, const int *__restrict__ r_ge,
           int shard, int n_shards, Acc1 *acc) {
    __shared__ __align__(16) float s_e[CT_TILE];
    __shared__ int s_red[2][32];
    __shared__ long long red[32];
    __shared__ unsigned long long s_item;
    const long long n_rows = (long long)acc->n_rows;
    const long long tiles_all = (n_rows + CT_ROWS - 1) / CT_ROWS;
    const long long my_tiles = tiles_all > shard ? (tiles_all - shard + n_shards - 1) / n_shards : 0;
    const long long col_tiles = (n + CT_TILE - 1) / CT_TILE;
    const unsigned long long total = (unsigned long long)(my_tiles * col_tiles);
    long long a = 0, b = 0, c = 0, d = 0;  // this CTA's strict conc / le and same-time conc / le counts
    for (;;) {
        __syncthreads();  // (s_item, s_e and s_red of the previous item are no longer read)
        if (threadIdx.x == 0) s_item = atomicAdd(&acc->work, 1ull);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= total) break;
        const long long ty = (long long)(item / (unsigned long long)col_tiles);
        const int c0 = (int)(item - (unsigned long long)ty * (unsigned long long)col_tiles) * CT_TILE;
        const int c1 = (int)min((long long)n, (long long)c0 + CT_TILE);
        const long long k0 = (ty * n_shards + shard) * CT_ROWS;  // row tile ty of this shard

        float lo[CT_R], hi[CT_R];
        int rs[CT_R], rg[CT_R];
        int mins = INT_MAX, maxge = 0;
#pragma unroll
        for (int u = 0; u < CT_R; ++u) {
            const long long k = k0 + threadIdx.x + (long long)u * CT_THREADS;
            if (k < n_rows) {
                lo[u] = r_lo[k]; hi[u] = r_hi[k]; rs[u] = r_s[k]; rg[u] = r_ge[k];
                mins = min(mins, rs[u]); maxge = max(maxge, rg[u]);
            } else {  // padding row: empty comparable range, NaN thresholds never compare true
                lo[u] = __int_as_float(0x7fc00000); hi[u] = lo[u]; rs[u] = INT_MAX; rg[u] = INT_MAX;
            }
        }
        // CTA-wide min(s) and max(ge) decide between skip / fast / general
        {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mins = min(mins, __shfl_xor_sync(FULL, mins, o));
                maxge = max(maxge, __shfl_xor_sync(FULL, maxge, o));
            }
            if (lane == 0) { s_red[0][wid] = mins; s_red[1][wid] = maxge; }
            __syncthreads();
            mins = INT_MAX; maxge = 0;
            for (int w = 0; w < CT_THREADS / 32; ++w) { mins = min(mins, s_red[0][w]); maxge = max(maxge, s_red[1][w]); }
        }
        if (c1 <= mins) continue;  // the whole column tile precedes every row's comparable range

        unsigned cs[CT_R], ls[CT_R];      // strict: #(e_j < lo), #(e_j <= hi)
        unsigned long long conc_t = 0, le_t = 0;  // same-time (rare, general path only)
#pragma unroll
        for (int u = 0; u < CT_R; ++u) { cs[u] = 0; ls[u] = 0; }
        for (int q = threadIdx.x; q < CT_TILE; q += CT_THREADS) s_e[q] = (c0 + q < n) ? est_s[c0 + q] : 0.f;
        __syncthreads();
        if (c0 >= maxge && c1 - c0 == CT_TILE) {
            // fast path: every (row, column) pair of this tile is a strict comparable pair
#pragma unroll 2
            for (int q = 0; q < CT_TILE; q += 4) {
                const float4 e4 = *reinterpret_cast<const float4 *>(s_e + q);
#pragma unroll
                for (int u = 0; u < CT_R; ++u) {
                    // compare + predicated increment (2 instructions per comparison; `count += (a < b)` compiles to
                    // ~2.5: predicate, select, add)
                    inc_lt(cs[u], e4.x, lo[u]); inc_lt(cs[u], e4.y, lo[u]); inc_lt(cs[u], e4.z, lo[u]); inc_lt(cs[u], e4.w, lo[u]);
                    inc_le(ls[u], e4.x, hi[u]); inc_le(ls[u], e4.y, hi[u]); inc_le(ls[u], e4.z, hi[u]); inc_le(ls[u], e4.w, hi[u]);
                }
            }
        } else {
            const int lim = c1 - c0;
            for (int q = 0; q < lim; ++q) {
                const float ej = s_e[q];
                const int j = c0 + q;
#pragma unroll
                for (int u = 0; u < CT_R; ++u) {
                    const bool lt = ej < lo[u], le = ej <= hi[u];
                    if (j >= rg[u]) { cs[u] += lt; ls[u] += le; }
                    else if (j >= rs[u]) { conc_t += lt; le_t += le; }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CT_R; ++u) { a += cs[u]; b += ls[u]; }
        c += (long long)conc_t; d += (long long)le_t;
    }
    a = block_reduce<long long>(a, 0ll, OpAddLL(), red);
    b = block_reduce<long long>(b, 0ll, OpAddLL(), red);
    c = block_reduce<long long>(c, 0ll, OpAddLL(), red);
    d = block_reduce<long long>(d, 0ll, OpAddLL(), red);
    if (threadIdx.x == 0) {
        if (a) atomicAdd(&acc->conc_s, (unsigned long long)a);
        if (b) atomicAdd(&acc->le_s, (unsigned long long)b);
        if (c) atomicAdd(&acc->conc_t, (unsigned long long)c);
        if (d) atomicAdd(&acc->le_t, (unsigned long long)d);
    }
}

// ------------------------------------------------------------------ algo 2: sorted tiles, ranks instead of pairs
// every FULL column tile sorted by estimate (ascending, NaN last) into est_t.  One CTA per tile, bitonic network on
// order-preserving keys in shared memory (-0 sorts before +0: a refinement of the float order, the searches' predicates
// treat them alike).
template <int N, bool PAYLOAD>
__device__ __forceinline__ void bitonic_sort_smem(uint32_t *s_k, float *s_p) {
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < N / 2; p += CT_THREADS) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i | j;   // the p-th pair (i, i ^ j) with i < l
                const uint32_t a = s_k[i], b = s_k[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    s_k[i] = b; s_k[l] = a;
                    if (PAYLOAD) { const float t = s_p[i]; s_p[i] = s_p[l]; s_p[l] = t; }
                }
            }
            __syncthreads();
        }
    }
}
__global__ void __launch_bounds__(CT_THREADS)
k_ci_tile_sort(const float *__restrict__ est_s, int64_t n, float *__restrict__ est_t) {
    __shared__ uint32_t s_k[CT_TILE];
    const int64_t c0 = (int64_t)blockIdx.x * CT_TILE;
    if (c0 + CT_TILE > n) return;   // the partial last tile is never searched
    for (int q = threadIdx.x; q < CT_TILE; q += CT_THREADS) {
        const float e = est_s[c0 + q];
        s_k[q] = (e == e) ? f2o(e) : 0xffffffffu;
    }
    __syncthreads();
    bitonic_sort_smem<CT_TILE, false>(s_k, nullptr);
    for (int q = threadIdx.x; q < CT_TILE; q += CT_THREADS) {
        const uint32_t o = s_k[q];
        est_t[c0 + q] = (o == 0xffffffffu) ? __int_as_float(0x7fc00000) : o2f(o);
    }
}
// every row tile: its rows' lo thresholds and its hi thresholds, each sorted on its own (NaN last, padded to the full tile with
// NaN: they never count) into r_lo2 / r_hi2 -- the sums over a row tile need neither the row a threshold belongs to nor the
// pairing of lo with hi -- and tile_rng[4 tile + {0, 1, 2, 3}] = min s, max ge, number of non-NaN lo, of non-NaN hi
__global__ void __launch_bounds__(CT_THREADS)
k_ci_row_sort(const float *__restrict__ r_lo, const float *__restrict__ r_hi, const int *__restrict__ r_s,
              const int *__restrict__ r_ge, const Acc1 *__restrict__ acc, float *__restrict__ r_lo2, float *__restrict__ r_hi2,
              int *__restrict__ tile_rng) {
    __shared__ uint32_t s_k[CT_ROWS], s_k2[CT_ROWS];
    __shared__ int s_red[4][32];
    const long long n_rows = (long long)acc->n_rows, k0 = (long long)blockIdx.x * CT_ROWS;
    if (k0 >= n_rows) return;
    int mins = INT_MAX, maxge = 0, nvl = 0, nvh = 0;
    for (int q = threadIdx.x; q < CT_ROWS; q += CT_THREADS) {
        const long long k = k0 + q;
        uint32_t kl = 0xffffffffu, kh = 0xffffffffu;
        if (k < n_rows) {
            const float lo = r_lo[k], hi = r_hi[k];
            if (lo == lo) { kl = f2o(lo); ++nvl; }
            if (hi == hi) { kh = f2o(hi); ++nvh; }
            mins = min(mins, r_s[k]); maxge = max(maxge, r_ge[k]);
        }
        s_k[q] = kl; s_k2[q] = kh;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mins = min(mins, __shfl_xor_sync(FULL, mins, o));
        maxge = max(maxge, __shfl_xor_sync(FULL, maxge, o));
        nvl += __shfl_xor_sync(FULL, nvl, o);
        nvh += __shfl_xor_sync(FULL, nvh, o);
    }
    if (lane == 0) { s_red[0][wid] = mins; s_red[1][wid] = maxge; s_red[2][wid] = nvl; s_red[3][wid] = nvh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        mins = INT_MAX; maxge = 0; nvl = 0; nvh = 0;
        for (int w = 0; w < CT_THREADS / 32; ++w) {
            mins = min(mins, s_red[0][w]); maxge = max(maxge, s_red[1][w]); nvl += s_red[2][w]; nvh += s_red[3][w];
        }
        int *o = tile_rng + 4 * blockIdx.x;
        o[0] = mins; o[1] = maxge; o[2] = nvl; o[3] = nvh;
    }
    bitonic_sort_smem<CT_ROWS, false>(s_k, nullptr);
    bitonic_sort_smem<CT_ROWS, false>(s_k2, nullptr);
    for (int q = threadIdx.x; q < CT_ROWS; q += CT_THREADS) {
        const uint32_t a = s_k[q], b = s_k2[q];
        r_lo2[k0 + q] = (a == 0xffffffffu) ? __int_as_float(0x7fc00000) : o2f(a);
        r_hi2[k0 + q] = (b == 0xffffffffu) ? __int_as_float(0x7fc00000) : o2f(b);
    }
}
// number of elements of the sorted tile s[0, 1024) (NaN last) with s[q] < v / s[q] <= v: ordered compares, so NaN elements
// and a NaN threshold count nothing -- exactly the sums the pair loop forms.  Branch-free bisection on the shared-window
// BYTE address: load with an immediate offset, compare, predicated add = three instructions per step (the C form of the
// same loop compiles to five: the select and the address add stay separate).
template <int STEP, bool LE>
__device__ __forceinline__ void tile_rank_step(uint32_t &a, float v) {
    float x;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(a), "n"((STEP - 1) * 4));
    if (LE) asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(a) : "f"(x), "f"(v), "n"(STEP * 4));
    else asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(a) : "f"(x), "f"(v), "n"(STEP * 4));
    if constexpr (STEP > 1) tile_rank_step<STEP / 2, LE>(a, v);
}
template <bool LE, int N = CT_TILE>
__device__ __forceinline__ unsigned tile_rank(uint32_t s_addr, float v) {   // s_addr: shared-window address of s[0]; N a power of 2
    uint32_t a = s_addr;
    tile_rank_step<N / 2, LE>(a, v);   // a <= s_addr + (N - 1) * 4
    tile_rank_step<1, LE>(a, v);       // the element the bisection ended on
    return (a - s_addr) >> 2;
}
__device__ __forceinline__ unsigned tile_rank_lt(const float *s, float v) { return tile_rank<false>(smem_addr_u32(s), v); }
__device__ __forceinline__ unsigned tile_rank_le(const float *s, float v) { return tile_rank<true>(smem_addr_u32(s), v); }

// the count kernel of algo 2: same work items as k_ci_count, pulled CI2_CHUNK consecutive items at a time (they share the row
// tile but for the chunk that crosses into the next one).
//   * column tile strictly later than every row of the row tile: the sums over the tile's rows are taken from the COLUMN side,
//       sum_rows #(e_j < lo_r) = sum_j #(lo_r > e_j) = sum_j (n_lo - rank_le(lo sorted, e_j)),
//       sum_rows #(e_j <= hi_r) = sum_j (n_hi - rank_lt(hi sorted, e_j))        (NaN columns count nothing),
//     1024 columns x 2 bisections of 12 steps against 2048 rows x 2 x 11 from the row side.  The row tile's two sorted
//     threshold arrays stay in shared memory while the items of a chunk share the row tile; a thread's four columns come
//     straight from the sorted column tile (neighbouring lanes hold neighbouring values: the bisections walk together).
//   * around the diagonal the rows keep their time order and their (lo, hi, s, ge): see below.
constexpr int CI2_CHUNK = 8;
__global__ void __launch_bounds__(CT_THREADS, 3)
k_ci_count2(const float *__restrict__ est_s, const float *__restrict__ est_t, int64_t n, const float *__restrict__ r_lo,
            const float *__restrict__ r_hi, const int *__restrict__ r_s, const int *__restrict__ r_ge,
            const float *__restrict__ r_lo2, const float *__restrict__ r_hi2, const int *__restrict__ tile_rng, int shard,
            int n_shards, Acc1 *acc) {
    __shared__ __align__(16) float s_e[CT_TILE];
    __shared__ __align__(16) float s_t[CT_TILE];
    __shared__ __align__(16) float s_lo[CT_ROWS], s_hi[CT_ROWS];
    __shared__ long long red[32];
    __shared__ unsigned long long s_item;
    __shared__ long long s_ty;
    const long long n_rows = (long long)acc->n_rows;
    const long long tiles_all = (n_rows + CT_ROWS - 1) / CT_ROWS;
    const long long my_tiles = tiles_all > shard ? (tiles_all - shard + n_shards - 1) / n_shards : 0;
    const long long col_tiles = (n + CT_TILE - 1) / CT_TILE;
    const unsigned long long total = (unsigned long long)(my_tiles * col_tiles);
    const uint32_t lo_addr = smem_addr_u32(s_lo), hi_addr = smem_addr_u32(s_hi);
    long long a = 0, b = 0, c = 0, d = 0;  // this CTA's strict conc / le and same-time conc / le counts
    long long cached_tg = -1;              // the row tile whose sorted thresholds sit in s_lo / s_hi
    // two passes over the items: first the few expensive ones around the diagonal, one at a time (balance), then the many
    // strictly-later ones in chunks that share their row tile's thresholds
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
    // pass 0 enumerates the items column tile by column tile (item = column tile * row tiles + row tile): the items around
    // the diagonal of one row tile are ~8 consecutive column tiles, so consecutive items of this order hold at most one or
    // two of them and a chunk of 32 stays balanced; pass 1 goes row tile by row tile (its chunk shares the thresholds)
    const unsigned long long per_cta = total / (4ull * gridDim.x);   // small cohorts: single items (balance before pull latency)
    const int cmax = pass == 0 ? 32 : CI2_CHUNK;
    const int chunk = (int)(per_cta < 1 ? 1 : (per_cta > (unsigned long long)cmax ? (unsigned long long)cmax : per_cta));
    unsigned long long *counter = pass == 0 ? &acc->work : &acc->work2;
    for (;;) {
        __syncthreads();  // (s_item of the previous chunk is no longer read)
        const long long inner = pass == 0 ? my_tiles : col_tiles;   // the fast index of this pass's enumeration
        if (threadIdx.x == 0) {
            const unsigned long long it = atomicAdd(counter, (unsigned long long)chunk);
            s_item = it;
            s_ty = inner > 0 ? (long long)(it / (unsigned long long)inner) : 0;   // one 64-bit division per chunk, not per item and thread
        }
        __syncthreads();
        const unsigned long long first = s_item;
        if (first >= total) break;
        long long outer = s_ty;
        long long in_ = (long long)(first - (unsigned long long)outer * (unsigned long long)inner);
        for (int gi = 0; gi < chunk; ++gi, ++in_) {
            const unsigned long long item = first + (unsigned long long)gi;
            if (item >= total) break;
            while (in_ >= inner) { in_ -= inner; ++outer; }
            const long long ty = pass == 0 ? in_ : outer, ct = pass == 0 ? outer : in_;
            const int c0 = (int)ct * CT_TILE;
            const int c1 = (int)min((long long)n, (long long)c0 + CT_TILE);
            const long long tg = ty * n_shards + shard, k0 = tg * CT_ROWS;  // row tile ty of this shard
            const int4 rng = *reinterpret_cast<const int4 *>(tile_rng + 4 * tg);   // min s, max ge, valid lo, valid hi
            if (c1 <= rng.x) continue;   // the column tile precedes every row's comparable range (block-uniform)
            const bool full = c1 - c0 == CT_TILE;
            const bool strictly_later = full && c0 >= rng.y;
            // pass 0: the full tiles around the diagonal; pass 1: the strictly-later tiles and the partial last column tile
            // (one per row tile: in pass 0's order they would all land in one chunk)
            if ((full && !strictly_later) != (pass == 0)) continue;
            if (strictly_later) {
                // every (row, column) pair of this item is a strict comparable pair
                float e[CT_TILE / CT_THREADS];
#pragma unroll
                for (int i = 0; i < CT_TILE / CT_THREADS; ++i) e[i] = est_t[c0 + threadIdx.x + i * CT_THREADS];
                if (cached_tg != tg) {
                    __syncthreads();   // the previous row tile's thresholds are no longer searched
                    for (int q = threadIdx.x; q < CT_ROWS / 4; q += CT_THREADS) {
                        reinterpret_cast<float4 *>(s_lo)[q] = reinterpret_cast<const float4 *>(r_lo2 + k0)[q];
                        reinterpret_cast<float4 *>(s_hi)[q] = reinterpret_cast<const float4 *>(r_hi2 + k0)[q];
                    }
                    cached_tg = tg;
                    __syncthreads();
                }
                unsigned sa = 0, sb = 0;
#pragma unroll
                for (int i = 0; i < CT_TILE / CT_THREADS; ++i) {
                    const unsigned ra = tile_rank<true, CT_ROWS>(lo_addr, e[i]), rb = tile_rank<false, CT_ROWS>(hi_addr, e[i]);
                    const bool ok = e[i] == e[i];
                    sa += ok ? (unsigned)rng.z - ra : 0u;
                    sb += ok ? (unsigned)rng.w - rb : 0u;
                }
                a += sa; b += sb;
                continue;
            }
            float lo[CT_R], hi[CT_R];
            int rs[CT_R], rg[CT_R];
#pragma unroll
            for (int u = 0; u < CT_R; ++u) {
                const long long k = k0 + threadIdx.x + (long long)u * CT_THREADS;
                if (k < n_rows) {
                    lo[u] = r_lo[k]; hi[u] = r_hi[k]; rs[u] = r_s[k]; rg[u] = r_ge[k];
                } else {  // padding row: empty comparable range, NaN thresholds never compare true
                    lo[u] = __int_as_float(0x7fc00000); hi[u] = lo[u]; rs[u] = INT_MAX; rg[u] = INT_MAX;
                }
            }
            __syncthreads();   // (s_e and s_t of the previous diagonal item are no longer read)
            for (int q = threadIdx.x; q < CT_TILE; q += CT_THREADS) s_e[q] = (c0 + q < n) ? est_s[c0 + q] : 0.f;
            if (full)
                for (int q = threadIdx.x; q < CT_TILE; q += CT_THREADS) s_t[q] = est_t[c0 + q];
            __syncthreads();
            unsigned long long conc_t = 0, le_t = 0;
            if (full) {
                // Around the diagonal.  Per row the tile splits into columns before its comparable range [0, qt), same-time
                // censored columns [qt, qs) and strictly later columns [qs, 1024): strict counts = ranks of the thresholds in
                // the whole sorted tile minus the counts over [0, qs); only the columns before qs are visited one by one.
#pragma unroll
                for (int u = 0; u < CT_R; ++u) {
                    const int qs = min(max(rg[u] - c0, 0), CT_TILE), qt = min(max(rs[u] - c0, 0), CT_TILE);   // padding rows: both 1024
                    unsigned pre_lt = 0, pre_le = 0, st_lt = 0, st_le = 0;
                    const bool strict_here = qs < CT_TILE;
                    if (strict_here)   // columns before the row's comparable range: only taken off the whole-tile ranks
                        for (int q = 0; q < qt; ++q) { const float ej = s_e[q]; inc_lt(pre_lt, ej, lo[u]); inc_le(pre_le, ej, hi[u]); }
                    for (int q = qt; q < qs; ++q) { const float ej = s_e[q]; inc_lt(st_lt, ej, lo[u]); inc_le(st_le, ej, hi[u]); }
                    if (strict_here) {
                        a += tile_rank_lt(s_t, lo[u]) - (pre_lt + st_lt); b += tile_rank_le(s_t, hi[u]) - (pre_le + st_le);
                    }
                    conc_t += st_lt; le_t += st_le;
                }
            } else {   // the partial last column tile: pair by pair
                const int lim = c1 - c0;
                unsigned cs[CT_R], ls[CT_R];
#pragma unroll
                for (int u = 0; u < CT_R; ++u) { cs[u] = 0; ls[u] = 0; }
                for (int q = 0; q < lim; ++q) {
                    const float ej = s_e[q];
                    const int j = c0 + q;
#pragma unroll
                    for (int u = 0; u < CT_R; ++u) {
                        const bool lt = ej < lo[u], le = ej <= hi[u];
                        if (j >= rg[u]) { cs[u] += lt; ls[u] += le; }
                        else if (j >= rs[u]) { conc_t += lt; le_t += le; }
                    }
                }
#pragma unroll
                for (int u = 0; u < CT_R; ++u) { a += cs[u]; b += ls[u]; }
            }
            c += (long long)conc_t; d += (long long)le_t;
        }
    }
    }
    a = block_reduce<long long>(a, 0ll, OpAddLL(), red);
    b = block_reduce<long long>(b, 0ll, OpAddLL(), red);
    c = block_reduce<long long>(c, 0ll, OpAddLL(), red);
    d = block_reduce<long long>(d, 0ll, OpAddLL(), red);
    if (threadIdx.x == 0) {
        if (a) atomicAdd(&acc->conc_s, (unsigned long long)a);
        if (b) atomicAdd(&acc->le_s, (unsigned long long)b);
        if (c) atomicAdd(&acc->conc_t, (unsigned long long)c);
        if (d) atomicAdd(&acc->le_t, (unsigned long long)d);
    }
}

__global__ void k_ci_final(const Acc1 *acc, long long *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] += (long long)acc->conc_s;
        out[1] += (long long)(acc->tot_s - acc->le_s);
        out[2] += (long long)(acc->le_s - acc->conc_s);
        out[3] += (long long)acc->conc_t;
        out[4] += (long long)(acc->tot_t - acc->le_t);
        out[5] += (long long)(acc->le_t - acc->conc_t);
    }
}

struct LoadIsRow {
    const int *isrow;
    __device__ sortscan::Tup operator()(int64_t p) const { sortscan::Tup t; t.a = 0.0; t.b = 0.0; t.i = isrow[p]; return t; }
};
struct StoreRank {  // exclusive prefix = index of the row among the selected event rows
    int *rank;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &el) const { rank[p] = (int)(inc.i - el.i); }
};

struct CiLayout {
    size_t off_acc, off_keys, off_vals, off_keys_s, off_idx_s, off_est_s, off_isrow, off_rank, off_lo, off_hi,
        off_s, off_ge, off_cub, cub_bytes, off_est_t, off_lo2, off_hi2, off_rng, total;
};
CiLayout ci_layout(int64_t n) {
    CiLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.off_acc = take(sizeof(Acc1));
    L.off_keys = take(N * 4); L.off_vals = take(N * 4); L.off_keys_s = take(N * 4); L.off_idx_s = take(N * 4);
    L.off_est_s = take(N * 4 + 16); L.off_isrow = take(N * 4); L.off_rank = take(N * 4);
    L.off_lo = take(N * 4); L.off_hi = take(N * 4); L.off_s = take(N * 4); L.off_ge = take(N * 4);
    const size_t tmp = sortscan::radix_sort_temp_bytes((int64_t)N), sc = sortscan::scan_state_bytes((int64_t)N);
    L.cub_bytes = (tmp > sc ? tmp : sc) + 256;
    L.off_cub = take(L.cub_bytes);
    // algo 2: the column tiles sorted by estimate, the row tiles' sorted thresholds (padded to whole tiles), [min s, max ge) per row tile
    L.off_est_t = take(N * 4 + 16); L.off_lo2 = take((N + CT_ROWS) * 4); L.off_hi2 = take((N + CT_ROWS) * 4);
    L.off_rng = take((N / CT_ROWS + 2) * 16);
    L.total = o;
    return L;
}

}  // namespace

size_t cindex_workspace_bytes(int64_t n, int algo) { return algo == 0 ? 256 : ci_layout(n).total; }

int32_t cindex_counts_launch(const float *est, const float *time, const uint8_t *event, int64_t n,
                             int64_t row_begin, int64_t row_end, float tol, int algo, int shard, int n_shards,
                             int64_t *out, void *ws, size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(n >= 0 && n < (int64_t)INT_MAX, "n must be < 2^31");
    B200_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n, "row range");
    B200_REQUIRE(tol >= 0.f, "tied_tol must be >= 0");
    B200_REQUIRE(n_shards >= 1 && shard >= 0 && shard < n_shards, "shard in [0, n_shards)");
    B200_REQUIRE(algo == 1 || algo == 2 || n_shards == 1, "tile shards need algo 1 or 2");
    if (n == 0 || row_begin == row_end) return B200SURV_OK;
    if (algo == 0) {
        const int64_t rows = row_end - row_begin;
        const unsigned gx = (unsigned)((rows + A0_THREADS - 1) / A0_THREADS);
        const unsigned gy = (unsigned)((n + A0_TILE * 16 - 1) / (A0_TILE * 16));
        B200_REQUIRE(gy <= 65535, "algo 0 supports n <= 2^30");
        cindex_direct<<<dim3(gx, gy), A0_THREADS, 0, st>>>(est, time, event, n, row_begin, row_end, tol,
                                                           reinterpret_cast<unsigned long long *>(out));
        B200_CHECK_CUDA(cudaGetLastError());
        return B200SURV_OK;
    }
    B200_REQUIRE(algo == 1 || algo == 2, "algo must be 0, 1 or 2");
    const CiLayout L = ci_layout(n);
    if (ws_bytes < L.total) { set_error("cindex: workspace %zu < %zu", ws_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    Acc1 *acc = reinterpret_cast<Acc1 *>(w8 + L.off_acc);
    uint32_t *keys = reinterpret_cast<uint32_t *>(w8 + L.off_keys), *vals = reinterpret_cast<uint32_t *>(w8 + L.off_vals),
             *keys_s = reinterpret_cast<uint32_t *>(w8 + L.off_keys_s), *idx_s = reinterpret_cast<uint32_t *>(w8 + L.off_idx_s);
    float *est_s = reinterpret_cast<float *>(w8 + L.off_est_s);
    int *isrow = reinterpret_cast<int *>(w8 + L.off_isrow), *rank = reinterpret_cast<int *>(w8 + L.off_rank);
    float *r_lo = reinterpret_cast<float *>(w8 + L.off_lo), *r_hi = reinterpret_cast<float *>(w8 + L.off_hi);
    int *r_s = reinterpret_cast<int *>(w8 + L.off_s), *r_ge = reinterpret_cast<int *>(w8 + L.off_ge);
    void *cub_tmp = w8 + L.off_cub;
    int grid = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    if (grid > cap) grid = cap;
    // keys are generated into (keys_s, idx_s): four ping-pong passes leave the sorted pairs there
    k_ci_keys<<<grid, 256, 0, st>>>(time, event, n, keys_s, idx_s, acc);
    {
        int in_first = 1;  // four passes: the sorted pairs are back in (keys_s, idx_s)
        const int32_t rc = sortscan::radix_sort_pairs2(keys_s, idx_s, keys, vals, n, 32, nullptr, 0, cub_tmp, st, &in_first);
        if (rc) return rc;
    }
    k_ci_flag<<<grid, 256, 0, st>>>(est, keys_s, idx_s, n, row_begin, row_end, est_s, isrow);
    {
        const int32_t rc = sortscan::scan_lookback<sortscan::I_ADD, false>(n, LoadIsRow{isrow}, StoreRank{rank}, cub_tmp, st);
        if (rc) return rc;
    }
    k_ci_rows<<<grid, 256, 0, st>>>(keys_s, est_s, isrow, rank, n, tol, r_lo, r_hi, r_s, r_ge, shard, n_shards, acc);
    if (algo == 2) {
        float *est_t = reinterpret_cast<float *>(w8 + L.off_est_t);
        float *r_lo2 = reinterpret_cast<float *>(w8 + L.off_lo2), *r_hi2 = reinterpret_cast<float *>(w8 + L.off_hi2);
        int *tile_rng = reinterpret_cast<int *>(w8 + L.off_rng);
        k_ci_tile_sort<<<(unsigned)((n + CT_TILE - 1) / CT_TILE), CT_THREADS, 0, st>>>(est_s, n, est_t);
        k_ci_row_sort<<<(unsigned)((n + CT_ROWS - 1) / CT_ROWS), CT_THREADS, 0, st>>>(r_lo, r_hi, r_s, r_ge, acc, r_lo2, r_hi2, tile_rng);
        k_ci_count2<<<8 * num_sms(), CT_THREADS, 0, st>>>(est_s, est_t, n, r_lo, r_hi, r_s, r_ge, r_lo2, r_hi2, tile_rng, shard, n_shards, acc);
    } else {
        k_ci_count<<<8 * num_sms(), CT_THREADS, 0, st>>>(est_s, n, r_lo, r_hi, r_s, r_ge, shard, n_shards, acc);
    }
    k_ci_final<<<1, 32, 0, st>>>(acc, reinterpret_cast<long long *>(out));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

// ---- test hooks for the primitives of sortscan.cuh (tests/test_sortscan_gpu.py)
size_t debug_sortscan_temp_bytes(int64_t n) {
    const size_t a = sortscan::radix_sort_temp_bytes(n), b = sortscan::scan_state_bytes(n);
    return (a > b ? a : b) + 256;
}
int32_t debug_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, int64_t n, void *temp,
                         cudaStream_t st) {
    int in_first = 1;
    return sortscan::radix_sort_pairs2(keys, vals, keys_tmp, vals_tmp, n, 32, nullptr, 0, temp, st, &in_first);  // sorted pairs end up in keys/vals
}
namespace {
struct DbgLoad {
    const double *a; const long long *i;
    __device__ sortscan::Tup operator()(int64_t p) const { sortscan::Tup t; t.a = a[p]; t.b = 2.0 * a[p]; t.i = i[p]; return t; }
};
struct DbgStore {
    double *a, *b; long long *i;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &) const { a[p] = inc.a; b[p] = inc.b; i[p] = inc.i; }
};
}  // namespace
// inclusive scan of (a, 2a, i) with i combined by iop (0 add, 1 min, 2 max), forward or reverse
int32_t debug_scan(const double *a, const long long *i, int64_t n, int iop, int reverse, double *out_a, double *out_b,
                   long long *out_i, void *temp, cudaStream_t st) {
    const DbgLoad ld{a, i};
    const DbgStore sto{out_a, out_b, out_i};
    if (iop == 0) return reverse ? sortscan::scan_lookback<sortscan::I_ADD, true>(n, ld, sto, temp, st)
                                 : sortscan::scan_lookback<sortscan::I_ADD, false>(n, ld, sto, temp, st);
    if (iop == 1) return reverse ? sortscan::scan_lookback<sortscan::I_MIN, true>(n, ld, sto, temp, st)
                                 : sortscan::scan_lookback<sortscan::I_MIN, false>(n, ld, sto, temp, st);
    return reverse ? sortscan::scan_lookback<sortscan::I_MAX, true>(n, ld, sto, temp, st)
                   : sortscan::scan_lookback<sortscan::I_MAX, false>(n, ld, sto, temp, st);
}

}  // namespace b200surv
