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
// Launch sequence (all hand-written):
//   keys    (time bits, censored bit); per-cohort max log_hz (the exponent shift), flags; the totals of the sort's first digit
//   sort    stable LSD radix sort (csrc/sortscan.cuh), 4 one-sweep passes of 8 bits on the key (+ 1-2 passes on the cohort
//           id for packed cohorts); the value of pair i is i, so the first pass reads no values
//   tiles   reduce-then-scan over tiles of 2048 sorted rows (see "tile kernels" below): the first sweep gathers
//           w = exp(log_hz - shift) through the permutation (kept as fp32 in sorted order), the second leaves per row what
//           of the gradient is known inside the tile, the third is element-wise and scatters the gradient through the
//           permutation; between them two scans over the per-tile records (one 8-CTA cluster per direction)
// Every group sum is a SEGMENTED sum of the group's own terms -- never a difference of two running totals: with hazards
// spread over tens of nats a late group's weights are 1e-15 of the total and a difference would be rounding noise.
#include <climits>
#include <cstdlib>

#include <cooperative_groups.h>

#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

using sortscan::Tup4;
namespace cg = cooperative_groups;

struct SegAcc {  // per cohort, device accumulators
    double sum_eta, sum_log;
    unsigned long long n_ev, n_times;
    float max_eta, max_time;
    unsigned flags;
    float neg_min_time;  // max of -time (single cohort only; time-range shards exchange it)
    double scale;  // d loss / d pll, written by k_tile_scan2 (shards: k_shard_finish)
};

// ---- time-range shards (multi-GPU, SURVEY.md 8e "Cox (B)"): shard r holds rows whose times are <= those of shard r + 1
// (ties across an edge allowed).  Every shard sorts and tiles its own rows; what crosses shards are three fixed-size
// records per shard, all-gathered by the caller, and the carries every shard folds out of them (k_shard_ctx*).
constexpr int SHARD_REC_BYTES = 128, SHARD_MAX = 64;
struct ShardRec0 {   // after the keys: what the neighbours and the common exponent shift need
    float max_eta, min_time, max_time;
    unsigned flags;
    long long n;
};
struct ShardRec1 {   // after the first tile scan: the shard's whole tile sequence folded into one element per chain
    double W;                       // total weight
    double RE, Rm; int Rflag;       // reverse chain of first fragments (event weight, events); flag: the shard holds a head
    int pad0;
    double LW, LE, Lm, Lrows; int Lflag, pad1;   // forward chain of last fragments
};
struct ShardRec2 {   // after the second tile scan
    double A, sum_log, sum_eta;     // sums of a, of log-denominators, of the event rows' log_hz
    long long n_times, n_ev;
    double AR, FR; int ARflag, pad0;   // reverse chain of first fragments (a, f)
    double FL; int FLflag; unsigned flags;
};
static_assert(sizeof(ShardRec0) <= SHARD_REC_BYTES && sizeof(ShardRec1) <= SHARD_REC_BYTES && sizeof(ShardRec2) <= SHARD_REC_BYTES,
              "shard records");
struct ShardCtx {    // device-resident, one per call: what the tile kernels of this shard add to their shard-local values
    uint32_t key_prev, key_next;    // time bits << 1 of the rows across the shard's edges
    int has_prev, has_next;
    double carS, carRE; int carRm, pad0;
    double carLW, carLE; int carLm, carLrows;
    double carC, carAR, carFR, carFL;
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
        a.max_eta = -INFINITY; a.max_time = -INFINITY; a.flags = 0; a.neg_min_time = -INFINITY; a.scale = 0.0;
        acc[s] = a;
    }
}

// keys, row indices, cohort ids (packed cohorts only), per-cohort max log_hz / max time / flags.  One cohort: per-thread
// running values and one atomic per block; packed cohorts: one atomic per warp and iteration when the warp's 32 consecutive
// rows share a cohort (else per lane).
__global__ void __launch_bounds__(256)
k_make_keys(const float *__restrict__ log_hz, const float *__restrict__ time, const uint8_t *__restrict__ event,
            const int64_t *__restrict__ seg_off, int n_seg, int64_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
            uint32_t *__restrict__ segid, SegAcc *acc, unsigned *__restrict__ hist0 /* the sort's first digit totals */) {
    __shared__ float red_f[32];
    __shared__ unsigned red_u[32];
    __shared__ unsigned s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    float mx = -INFINITY, mt = -INFINITY, mn = -INFINITY;
    unsigned flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t first = 0;   // rows [0, first) are done four at a time below
    if (seg_off == nullptr && ((reinterpret_cast<uintptr_t>(log_hz) | reinterpret_cast<uintptr_t>(time) | reinterpret_cast<uintptr_t>(keys)) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(event) & 3) == 0) {
        // one cohort, aligned inputs: four rows per thread and iteration (16-byte loads and stores; the scalar loop below issues
        // 65 instructions per row, most of them index arithmetic and bounds tests)
        const int64_t ngroups = n >> 2;
        for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
            const float4 t4 = __ldg(reinterpret_cast<const float4 *>(time) + g), e4 = __ldg(reinterpret_cast<const float4 *>(log_hz) + g);
            const uint32_t v4 = __ldg(reinterpret_cast<const uint32_t *>(event) + g);
            uint4 k4;
            k4.x = time_key(t4.x, (v4 & 0xffu) != 0); k4.y = time_key(t4.y, (v4 & 0xff00u) != 0);
            k4.z = time_key(t4.z, (v4 & 0xff0000u) != 0); k4.w = time_key(t4.w, (v4 & 0xff000000u) != 0);
            reinterpret_cast<uint4 *>(keys)[g] = k4;
            atomicAdd(&s_hist[k4.x & 255u], 1u); atomicAdd(&s_hist[k4.y & 255u], 1u);
            atomicAdd(&s_hist[k4.z & 255u], 1u); atomicAdd(&s_hist[k4.w & 255u], 1u);
            mx = fmaxf(fmaxf(mx, fmaxf(e4.x, e4.y)), fmaxf(e4.z, e4.w));
            const float tmax = fmaxf(fmaxf(t4.x, t4.y), fmaxf(t4.z, t4.w)), tmin = fminf(fminf(t4.x, t4.y), fminf(t4.z, t4.w));
            mt = fmaxf(mt, tmax); mn = fmaxf(mn, -tmin);
            if (!(t4.x >= 0.f) || !(t4.y >= 0.f) || !(t4.z >= 0.f) || !(t4.w >= 0.f)) flags |= B200SURV_COXF_BAD_TIME;
        }
        first = ngroups << 2;
    }
    const int64_t n_round = first + (n - first + 31) / 32 * 32;  // whole warps iterate together (the shuffles below need every lane)
    for (int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < n;
        const float t = in ? time[i] : 0.f, e = in ? log_hz[i] : -INFINITY;
        const unsigned bad = (in && !(t >= 0.f)) ? B200SURV_COXF_BAD_TIME : 0u;
        if (in) {   // the value of pair i is i: the sort's first pass does not read it
            const uint32_t key = time_key(t, event[i] != 0);
            keys[i] = key;
            atomicAdd(&s_hist[key & 255u], 1u);
        }
        if (seg_off == nullptr) {
            mx = fmaxf(mx, e); mt = fmaxf(mt, in ? t : -INFINITY); mn = fmaxf(mn, in ? -t : -INFINITY); flags |= bad;
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
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist0[threadIdx.x], s_hist[threadIdx.x]);
    if (seg_off == nullptr) {
        mx = block_reduce<float>(mx, -INFINITY, OpMaxF(), red_f);
        mt = block_reduce<float>(mt, -INFINITY, OpMaxF(), red_f);
        mn = block_reduce<float>(mn, -INFINITY, OpMaxF(), red_f);
        flags = block_reduce<unsigned>(flags, 0u, OpOrU(), red_u);
        if (threadIdx.x == 0) {
            atomic_max_float(&acc->max_eta, mx);
            atomic_max_float(&acc->max_time, mt);
            atomic_max_float(&acc->neg_min_time, mn);
            if (flags) atomicOr(&acc->flags, flags);
        }
    }
}

// ------------------------------------------------------------------------------------------------ tile kernels
// Risk-set sums, Efron terms, loss and gradient as REDUCE-THEN-SCAN over tiles of 2048 sorted rows: no look-back chain and
// no per-row fp64 intermediate in HBM.  A tile never crosses a cohort.  Tie groups that cross tile boundaries are handled
// through per-tile FRAGMENT records (rows before the tile's first group head / from its last head on) chained by the two
// tile scans (one cluster of 8 CTAs per direction):
//   k_tile_w      per tile: w gathered and stored; total weight, the two fragments' (weight, event weight, event count,
//                 rows); the cohort's sum of event log_hz and event count                [12 B/row + the gather, 4 B/row written]
//   k_tile_scan1  over tiles: S = weight of the cohort's later tiles; R = (E, m) of the rows AFTER the tile that belong to
//                 its last row's group; L = (W, E, m, rows) of the rows BEFORE the tile that belong to its first row's group
//   k_tile_terms  per tile: every group's (D, E, m) -> a_p, f_p, log-denominators; in-tile prefix P and group sums F; tile
//                 sums and fragment sums (read off the same forward scan); per row g0 = d - w (P - d F) at its group's end
//                 inside the tile and three flag bits                                            [8 B/row read, 5 B/row written]
//   k_tile_scan2  over tiles: C = sum of a over the cohort's earlier tiles; AR, FR / FL = the open groups' sums of a, f in
//                 later / earlier tiles; per-cohort loss, scale and header
//   k_tile_apply  element-wise: grad = scale (g0 - w (C + [open] AR) + w d ([open] FR + [began earlier] FL)), scattered
//                 through the permutation with L2 evict-last stores                              [13 B/row + the scatter]
// Inside a tile the work is three block-wide scans over a blocked arrangement (8 consecutive rows per thread): reverse
// (suffix weight; group-suffix event weight / count restarting at group tails; nearest tail), forward max (group start),
// forward (prefix of a; group-prefix of f restarting at heads); group values reach their rows through shared memory.
// The same tile sequence can span several GPUs (time-range shards): the tile records are then all-gathered and every rank
// runs the tile scans over the whole sequence (b200surv_cox_sorted_* phase entry points, dist.py).
constexpr int TS_THREADS = 256, TS_ITEMS = 8, TS_TILE = TS_THREADS * TS_ITEMS;
constexpr int TS_NW = TS_THREADS / 32;
constexpr int TF_FIRST = 1, TF_LAST = 2;   // first / last tile of its cohort
constexpr int TS_KN = TS_TILE + 2 + (TS_TILE + 2) / 32 + 2;   // skewed 4-byte array with a halo of two
constexpr int TS_DN = TS_TILE + TS_TILE / 8;                  // skewed 8-byte array

struct TileW {   // 64 bytes
    double Wtot, Ef, Wl, El;
    int mf, ml, rowsf, rowsl, nheads, flags, pad0, pad1;
};
struct TileC1 {  // 48 bytes
    double S, RE, LW, LE;
    int Rm, Lm, Lrows;
    int cf;   // shards: bit 0 = no head in the shard's later tiles (add the R carry), bit 1 = ... earlier tiles (add the L carry)
};
struct TileA {   // 64 bytes
    double sumA, sumL, Af, Ff, Al, Fl;
    int n_times, n_ev, pad0, pad1;
};
struct TileC2 { double C, AR, FR, FL; int cr, cfw; };   // cr / cfw: as TileC1::cf, for (AR, FR) / FL
static_assert(sizeof(TileW) == 64 && sizeof(TileA) == 64 && sizeof(TileC1) == 48 && sizeof(TileC2) == 40, "tile records");

struct TileGeo {
    int64_t p0;
    int rows, seg, flags;
    bool valid;
    int halo;                      // shards: bit 0 / 1 = the row before / after the tile lives on the previous / next shard
    uint32_t key_prev, key_next;
};
// tile -> (cohort, first row, rows).  One cohort: tiles of the whole range; packed cohorts: tile_base[s] = number of tiles
// of the cohorts before s (k_tile_base), every cohort starts a new tile.
__device__ __forceinline__ TileGeo tile_geo(int64_t t, const int64_t *__restrict__ seg_off, const int64_t *__restrict__ tile_base,
                                            int n_seg, int64_t n, const ShardCtx *__restrict__ ctx = nullptr) {
    TileGeo g;
    g.halo = 0; g.key_prev = 0u; g.key_next = 0u;
    if (seg_off == nullptr) {
        const int64_t nt = (n + TS_TILE - 1) / TS_TILE;
        g.valid = t < nt; g.p0 = t * TS_TILE; g.seg = 0;
        g.rows = (int)(n - g.p0 < TS_TILE ? n - g.p0 : TS_TILE);
        g.flags = (t == 0 ? TF_FIRST : 0) | (t == nt - 1 ? TF_LAST : 0);
        if (ctx != nullptr) {   // the cohort goes on across the shard's edges
            if (t == 0 && ctx->has_prev) { g.flags &= ~TF_FIRST; g.halo |= 1; g.key_prev = ctx->key_prev; }
            if (t == nt - 1 && ctx->has_next) { g.flags &= ~TF_LAST; g.halo |= 2; g.key_next = ctx->key_next; }
        }
        return g;
    }
    g.valid = t < tile_base[n_seg];
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {  // largest s with tile_base[s] <= t (an empty cohort shares its base with the next one)
        const int mid = (lo + hi + 1) >> 1;
        if (tile_base[mid] <= t) lo = mid; else hi = mid - 1;
    }
    g.seg = lo;
    const int64_t c1 = seg_off[lo + 1];
    g.p0 = seg_off[lo] + (t - tile_base[lo]) * TS_TILE;
    g.rows = g.valid ? (int)(c1 - g.p0 < TS_TILE ? c1 - g.p0 : TS_TILE) : 0;
    g.flags = (t == tile_base[lo] ? TF_FIRST : 0) | (t + 1 == tile_base[lo + 1] ? TF_LAST : 0);
    return g;
}

__global__ void __launch_bounds__(1024)
k_tile_base(const int64_t *__restrict__ seg_off, int n_seg, int64_t *__restrict__ tile_base) {
    __shared__ int64_t s_part[1024];
    const int t = threadIdx.x, per = (n_seg + 1023) / 1024;
    int64_t sum = 0;
    for (int k = 0; k < per; ++k) {
        const int s = t * per + k;
        if (s < n_seg) sum += (seg_off[s + 1] - seg_off[s] + TS_TILE - 1) / TS_TILE;
    }
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const int64_t v = s_part[i]; s_part[i] = run; run += v; }
    }
    __syncthreads();
    int64_t run = s_part[t];
    for (int k = 0; k < per; ++k) {
        const int s = t * per + k;
        if (s < n_seg) { tile_base[s] = run; run += (seg_off[s + 1] - seg_off[s] + TS_TILE - 1) / TS_TILE; }
        if (s == n_seg - 1) tile_base[n_seg] = run;
    }
}

// ---- block-wide scans over a blocked arrangement.  T: identity(), combine(first, second) with `first` earlier in SCAN
// order, shfl(lane).  REV: scan order = descending thread index.  Returns the combination of everything before this thread.
template <typename T, bool REV>
__device__ __forceinline__ T block_prefix(const T &agg, T *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = agg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T u = inc.shfl((REV ? lane + d : lane - d) & 31);
        if (REV ? (lane + d < 32) : (lane >= d)) inc = T::combine(u, inc);
    }
    __syncthreads();  // s_warp may still be read from the previous scan
    if (lane == (REV ? 0 : 31)) s_warp[warp] = inc;
    __syncthreads();
    // the warps' totals: every warp scans them itself, lane l holding the total at scan position l (three shuffle steps and one
    // broadcast instead of a loop of TS_NW - 1 combines out of shared memory in every thread)
    static_assert(TS_NW <= 32, "one lane per warp total");
    const int pos = REV ? TS_NW - 1 - warp : warp;
    T v = lane < TS_NW ? s_warp[REV ? TS_NW - 1 - lane : lane] : T::identity();
#pragma unroll
    for (int d = 1; d < TS_NW; d <<= 1) {
        const T u = v.shfl((lane - d) & 31);
        if (lane >= d) v = T::combine(u, v);
    }
    T pre = v.shfl((pos - 1) & 31);
    if (pos == 0) pre = T::identity();
    T ex = inc.shfl((REV ? lane + 1 : lane - 1) & 31);
    if (lane == (REV ? 31 : 0)) ex = T::identity();
    return T::combine(pre, ex);
}

struct RevT {   // reverse scan: W = suffix weight; E, m = group-suffix (restart at tails); tpos = nearest tail at or after
    double W, E;
    int m, tpos;
    static __device__ __forceinline__ RevT identity() { RevT r; r.W = 0.0; r.E = 0.0; r.m = 0; r.tpos = INT_MAX; return r; }
    // x: rows after y's rows (already accumulated in reverse order)
    static __device__ __forceinline__ RevT combine(const RevT &x, const RevT &y) {
        RevT r;
        const bool yt = y.tpos != INT_MAX;
        r.W = x.W + y.W; r.E = yt ? y.E : x.E + y.E; r.m = yt ? y.m : x.m + y.m; r.tpos = min(x.tpos, y.tpos);
        return r;
    }
    __device__ __forceinline__ RevT shfl(int src) const {
        RevT r;
        r.W = __shfl_sync(FULL, W, src); r.E = __shfl_sync(FULL, E, src); r.m = __shfl_sync(FULL, m, src);
        r.tpos = __shfl_sync(FULL, tpos, src);
        return r;
    }
};
struct MaxT {   // forward max scan of head positions (-1: none)
    int h;
    static __device__ __forceinline__ MaxT identity() { MaxT r; r.h = -1; return r; }
    static __device__ __forceinline__ MaxT combine(const MaxT &x, const MaxT &y) { MaxT r; r.h = max(x.h, y.h); return r; }
    __device__ __forceinline__ MaxT shfl(int src) const { MaxT r; r.h = __shfl_sync(FULL, h, src); return r; }
};
struct FwdT {   // forward scan: A = prefix of a; As, F = group-prefixes of a, f (restart at heads)
    double A, As, F;
    int head;
    static __device__ __forceinline__ FwdT identity() { FwdT r; r.A = 0.0; r.As = 0.0; r.F = 0.0; r.head = 0; return r; }
    static __device__ __forceinline__ FwdT combine(const FwdT &x, const FwdT &y) {
        FwdT r;
        r.A = x.A + y.A; r.As = y.head ? y.As : x.As + y.As; r.F = y.head ? y.F : x.F + y.F; r.head = x.head | y.head;
        return r;
    }
    __device__ __forceinline__ FwdT shfl(int src) const {
        FwdT r;
        r.A = __shfl_sync(FULL, A, src); r.As = __shfl_sync(FULL, As, src); r.F = __shfl_sync(FULL, F, src);
        r.head = __shfl_sync(FULL, head, src);
        return r;
    }
};

__device__ __forceinline__ int sk(int j) { return j + (j >> 3); }    // skewed index of an 8-byte array read with stride 8
__device__ __forceinline__ int sk4(int j) { return j + (j >> 5); }   // ... of a 4-byte array

// What a thread holds of its 8 rows after tile_load: weight and flags (bit 0 head, 1 tail, 2 event).
struct TileRows {
    float w[TS_ITEMS];
    unsigned flg[TS_ITEMS];
};
constexpr unsigned RF_HEAD = 1, RF_TAIL = 2, RF_EV = 4;

// keys and weights of the tile through shared memory (coalesced loads at any alignment), then the blocked rows' flags.
// A tile's first row continues the previous tile's group unless it is a head: equal time bits across the edge.  Rows past
// the end of a partial tile are neutral (weight 0, no flags).
// Where a tile's weights come from in the FIRST sweep: w = exp(log_hz - shift) gathered through the permutation and written
// out in sorted order for the later sweeps (the element-wise k_weights kernel of earlier versions, fused into the tile's
// loader: one launch, one read of the keys and one write + read of w less).  The cohort's sum of the event rows' log_hz and
// its event count ride along (eta_sum / n_ev, per thread).
struct WeightSrc {
    const float *log_hz;
    const uint32_t *idx_s;
    float *w_out;
    float shift;
};
template <bool STREAM = false, bool GATHER = false>
__device__ __forceinline__ void tile_load(const TileGeo &g, const uint32_t *__restrict__ keys_s, const float *__restrict__ w,
                                          uint32_t *s_key /*[TS_KN]*/, float *s_w /*[TS_KN]*/, TileRows &R,
                                          const WeightSrc *src = nullptr, double *eta_sum = nullptr, int *n_ev = nullptr) {
    const int t = threadIdx.x;
    if (!GATHER && !STREAM && g.rows == TS_TILE && !(g.flags & (TF_FIRST | TF_LAST)) && g.halo == 0) {
        // a whole tile inside its cohort (all but the first, the last and the partial ones): no bounds or edge tests
        const uint32_t *kp = keys_s + (g.p0 - 1) + t;
        const float *wp = w + g.p0 + t;
        uint32_t kq[TS_ITEMS + 1];
        float wq[TS_ITEMS];
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) { kq[k] = kp[k * TS_THREADS]; wq[k] = wp[k * TS_THREADS]; }
        kq[TS_ITEMS] = t < 2 ? kp[TS_ITEMS * TS_THREADS] : 0u;
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t + k * TS_THREADS;
            s_key[j + (j >> 5)] = kq[k]; s_w[j + (j >> 5)] = wq[k];
        }
        if (t < 2) s_key[t + TS_TILE + ((t + TS_TILE) >> 5)] = kq[TS_ITEMS];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t * TS_ITEMS + k;
            const uint32_t a = s_key[j + (j >> 5)], b = s_key[j + 1 + ((j + 1) >> 5)], c = s_key[j + 2 + ((j + 2) >> 5)];
            R.w[k] = s_w[j + (j >> 5)];
            R.flg[k] = ((a >> 1) != (b >> 1) ? RF_HEAD : 0u) | ((c >> 1) != (b >> 1) ? RF_TAIL : 0u) | ((b & 1u) ? 0u : RF_EV);
        }
        return;
    }
    const uint64_t pol = STREAM ? l2_policy_evict_first() : 0ull;   // last reader of (keys, w): do not displace the scatter target
    // s_key[1 + j] = key of row j; s_key[0] / s_key[rows + 1] = the neighbours across the tile edges.  All loads of a thread
    // are issued before the first store (one memory latency per tile, not one per element).
    uint32_t kk[TS_ITEMS + 1];
    float ww[TS_ITEMS];
#pragma unroll
    for (int k = 0; k < TS_ITEMS + 1; ++k) {
        const int j = t + k * TS_THREADS;
        const bool edge = (j == 0 && ((g.flags & TF_FIRST) || (g.halo & 1))) || (j == g.rows + 1 && ((g.flags & TF_LAST) || (g.halo & 2)));
        kk[k] = (j < g.rows + 2 && !edge) ? (STREAM ? ldg_hint_u32(keys_s + g.p0 + j - 1, pol) : keys_s[g.p0 + j - 1]) : 0u;
        if (j == 0 && (g.halo & 1)) kk[k] = g.key_prev;
        if (j == g.rows + 1 && (g.halo & 2)) kk[k] = g.key_next;
    }
    if (GATHER) {
        const uint64_t pol_s = l2_policy_evict_first(), pol_k = l2_policy_evict_last();   // the gather over log_hz stays in L2
        uint32_t ri[TS_ITEMS], kc[TS_ITEMS];
        float eta[TS_ITEMS];
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t + k * TS_THREADS;
            ri[k] = j < g.rows ? ldg_hint_u32(src->idx_s + g.p0 + j, pol_s) : 0u;
            kc[k] = j < g.rows ? keys_s[g.p0 + j] : 1u;
        }
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t + k * TS_THREADS;
            eta[k] = j < g.rows ? ldg_hint_f32(src->log_hz + ri[k], pol_k) : 0.f;
        }
        double se = 0.0;
        int ne = 0;
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t + k * TS_THREADS;
            // fp32: the difference of two floats rounds once, expf is good to 1 ulp, and w is kept as fp32 anyway
            ww[k] = j < g.rows ? expf(eta[k] - src->shift) : 0.f;
            if (j < g.rows) src->w_out[g.p0 + j] = ww[k];
            if (!(kc[k] & 1u)) { se += (double)eta[k]; ++ne; }
        }
        *eta_sum = se; *n_ev = ne;
    } else {
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            const int j = t + k * TS_THREADS;
            ww[k] = j < g.rows ? (STREAM ? ldg_hint_f32(w + g.p0 + j, pol) : w[g.p0 + j]) : 0.f;
        }
    }
#pragma unroll
    for (int k = 0; k < TS_ITEMS + 1; ++k) {
        const int j = t + k * TS_THREADS;
        if (j < g.rows + 2) s_key[j + (j >> 5)] = kk[k];
    }
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t + k * TS_THREADS;
        if (j < g.rows) s_w[j + (j >> 5)] = ww[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t * TS_ITEMS + k;
        if (j < g.rows) {
            const int j0 = j, j1 = j + 1, j2 = j + 2;
            const uint32_t kp = s_key[j0 + (j0 >> 5)], kc = s_key[j1 + (j1 >> 5)], kn = s_key[j2 + (j2 >> 5)];
            R.w[k] = s_w[j + (j >> 5)];
            const bool head = (j == 0 && (g.flags & TF_FIRST)) || (kp >> 1) != (kc >> 1);
            const bool tail = (j == g.rows - 1 && (g.flags & TF_LAST)) || (kn >> 1) != (kc >> 1);
            R.flg[k] = (head ? RF_HEAD : 0u) | (tail ? RF_TAIL : 0u) | ((kc & 1u) ? 0u : RF_EV);
        } else {
            R.w[k] = 0.f; R.flg[k] = 0u;
        }
    }
}

// 1 / x and log x of a positive normal double from fp32 seeds on the mantissa (x = m 2^e, m in [1, 2)): two Newton steps
// give the reciprocal to fp64 rounding; the logarithm is e ln 2 + logf(m), absolute error < 1e-7 on terms of size ~10 that
// are averaged over the events (the full fp64 routines were half of the instructions of these kernels)
__device__ __forceinline__ double fast_rcp(double x) {
    const int hi = __double2hiint(x);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    double r = (double)__frcp_rn((float)m);
    r = r * (2.0 - m * r);
    r = r * (2.0 - m * r);
    return __hiloint2double(__double2hiint(r) - (e << 20), __double2loint(r));
}
__device__ __forceinline__ double fast_log(double x) {
    const int hi = __double2hiint(x);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    return (double)e * 0.6931471805599453 + (double)logf((float)m);
}

// first sweep: the tile's total weight and its two fragments, as masked sums (no scan needed yet)
__global__ void __launch_bounds__(TS_THREADS, 5)
k_tile_w(const float *__restrict__ log_hz, const uint32_t *__restrict__ keys_s, const uint32_t *__restrict__ idx_s,
         float *__restrict__ w, const int64_t *__restrict__ seg_off, const int64_t *__restrict__ tile_base, int n_seg, int64_t n,
         const ShardCtx *__restrict__ ctx, SegAcc *acc, TileW *__restrict__ tw) {
    __shared__ uint32_t s_key[TS_KN];
    __shared__ float s_w[TS_KN];
    __shared__ int s_hf[TS_NW], s_hl[TS_NW], s_nh[TS_NW];
    __shared__ double s_sum[5][TS_NW];
    __shared__ int s_cnt[3][TS_NW];
    const TileGeo g = tile_geo(blockIdx.x, seg_off, tile_base, n_seg, n, ctx);
    if (!g.valid) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    TileRows R;
    WeightSrc src;
    src.log_hz = log_hz; src.idx_s = idx_s; src.w_out = w; src.shift = acc[g.seg].max_eta;
    double se;
    int ne;
    tile_load<false, true>(g, keys_s, nullptr, s_key, s_w, R, &src, &se, &ne);
    int hf = INT_MAX, hl = -1, nh = 0;
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k)
        if (R.flg[k] & RF_HEAD) { const int j = t * TS_ITEMS + k; hf = min(hf, j); hl = max(hl, j); ++nh; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hf = min(hf, __shfl_xor_sync(FULL, hf, o)); hl = max(hl, __shfl_xor_sync(FULL, hl, o)); nh += __shfl_xor_sync(FULL, nh, o);
    }
    if (lane == 0) { s_hf[warp] = hf; s_hl[warp] = hl; s_nh[warp] = nh; }
    __syncthreads();
    int first = INT_MAX, last = -1, nheads = 0;
#pragma unroll
    for (int q = 0; q < TS_NW; ++q) { first = min(first, s_hf[q]); last = max(last, s_hl[q]); nheads += s_nh[q]; }
    if (!nheads) { first = g.rows; last = 0; }   // no head: both fragments are the whole tile
    double v[5] = {0.0, 0.0, 0.0, 0.0, se};      // Wtot, Ef, Wl, El; the event rows' log_hz
    int mf = 0, ml = 0;
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t * TS_ITEMS + k;
        const double wk = (double)R.w[k], ek = (R.flg[k] & RF_EV) ? wk : 0.0;
        const int dk = (R.flg[k] & RF_EV) ? 1 : 0;
        v[0] += wk;
        if (j < first) { v[1] += ek; mf += dk; }
        if (j >= last) { v[2] += wk; v[3] += ek; ml += dk; }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = warp_sum(v[i]);
    mf = (int)warp_sum((long long)mf); ml = (int)warp_sum((long long)ml); ne = (int)warp_sum((long long)ne);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) s_sum[i][warp] = v[i];
        s_cnt[0][warp] = mf; s_cnt[1][warp] = ml; s_cnt[2][warp] = ne;
    }
    __syncthreads();
    if (t == 0) {
        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        int cf = 0, cl = 0, ce = 0;
        for (int q = 0; q < TS_NW; ++q) {
            for (int i = 0; i < 5; ++i) r[i] += s_sum[i][q];
            cf += s_cnt[0][q]; cl += s_cnt[1][q]; ce += s_cnt[2][q];
        }
        if (ce > 0) { atomicAdd(&acc[g.seg].sum_eta, r[4]); atomicAdd(&acc[g.seg].n_ev, (unsigned long long)ce); }
        TileW o;
        o.Wtot = r[0]; o.Ef = r[1]; o.Wl = r[2]; o.El = r[3]; o.mf = cf; o.ml = cl;
        o.rowsf = first; o.rowsl = g.rows - last; o.nheads = nheads; o.flags = g.flags; o.pad0 = 0; o.pad1 = 0;
        tw[blockIdx.x] = o;
    }
}

// ---- the two scans over tiles (one CTA).  Chains of fragments are segmented sums: they restart at every tile that holds
// a head; the sums over a cohort's tiles restart at the cohort's first / last tile.
template <int N>
struct SegN {
    double v[N];
    int flag;
    static __device__ __forceinline__ SegN identity() { SegN r; for (int i = 0; i < N; ++i) r.v[i] = 0.0; r.flag = 0; return r; }
    static __device__ __forceinline__ SegN pick(bool take, const SegN &x) {            // x or the identity, without a branch
        SegN r;
        for (int i = 0; i < N; ++i) r.v[i] = take ? x.v[i] : 0.0;
        r.flag = take ? x.flag : 0;
        return r;
    }
    static __device__ __forceinline__ SegN combine(const SegN &x, const SegN &y) {   // x before y in scan order
        SegN r;
        for (int i = 0; i < N; ++i) r.v[i] = y.flag ? y.v[i] : x.v[i] + y.v[i];
        r.flag = x.flag | y.flag;
        return r;
    }
};
constexpr int SC_THREADS = 512;
template <int N>
__device__ __forceinline__ SegN<N> seg_shfl(const SegN<N> &v, int src) {
    SegN<N> r;
    for (int i = 0; i < N; ++i) r.v[i] = __shfl_sync(FULL, v.v[i], src);
    r.flag = __shfl_sync(FULL, v.flag, src);
    return r;
}
// One round of a tile scan: thread i holds the element at scan position i of the round; returns the EXCLUSIVE prefix inside
// the round and leaves the round's total in s_warp[32] (valid after the call for every thread).
template <int N>
__device__ __forceinline__ SegN<N> round_prefix(const SegN<N> &e, SegN<N> *s_warp /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // branch-free steps (a lane without a partner combines with the identity): a shuffle inside divergent code takes the
    // WARPSYNC.COLLECTIVE slow path on this part
    __syncwarp();
    SegN<N> inc = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const SegN<N> u = seg_shfl(inc, (lane - d) & 31);
        inc = SegN<N>::combine(SegN<N>::pick(lane >= d, u), inc);
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        SegN<N> v = lane < SC_THREADS / 32 ? s_warp[lane] : SegN<N>::identity();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const SegN<N> u = seg_shfl(v, (lane - d) & 31);
            v = SegN<N>::combine(SegN<N>::pick(lane >= d, u), v);
        }
        s_warp[lane] = v;
        if (lane == SC_THREADS / 32 - 1) s_warp[32] = v;   // the round's total (lanes beyond the warp count hold nothing)
    }
    __syncthreads();
    const SegN<N> pre = warp ? s_warp[warp - 1] : SegN<N>::identity();
    SegN<N> ex = seg_shfl(inc, (lane - 1) & 31);
    if (lane == 0) ex = SegN<N>::identity();
    return SegN<N>::combine(pre, ex);
}

// A round covers SC_THREADS * SC_ITEMS tiles: every thread folds SC_ITEMS consecutive tiles (in scan order) on its own, one
// block-wide prefix per round joins the threads, a carry joins the rounds.  pos = position in scan order; a reverse scan
// visits tile T - 1 - pos.
// One CTA per direction was bound by what ONE SM can keep in flight (512 KB of tile records per scan at 16.7M rows: 50 us,
// ncu r2_v3).  Each direction is now a CLUSTER of SC_CLUSTER CTAs: CTA k of the cluster takes a contiguous run of rounds,
// walks it twice -- first for its total only, then, after the totals have met through distributed shared memory, with the
// fold of the earlier CTAs' totals as carry-in (the second walk reads from L1 / L2).
constexpr int SC_ITEMS = 2, SC_ROUND = SC_THREADS * SC_ITEMS, SC_CLUSTER = 8;

struct ScanRange { int64_t r0, r1; };
__device__ __forceinline__ ScanRange scan_range(int64_t T, unsigned rank) {
    const int64_t rounds = (T + SC_ROUND - 1) / SC_ROUND, per = (rounds + SC_CLUSTER - 1) / SC_CLUSTER;
    ScanRange g;
    g.r0 = (int64_t)rank * per < rounds ? (int64_t)rank * per : rounds;
    g.r1 = g.r0 + per < rounds ? g.r0 + per : rounds;
    return g;
}
// total of this CTA's rounds -> fold of the totals of the cluster's earlier CTAs (identical in every thread).
// s_x[0]: this CTA's total, read by the later CTAs; s_x[1]: the fold.  The caller ends with one more cluster.sync().
template <typename S>
__device__ __forceinline__ S cluster_carry(cg::cluster_group &cl, const S &total, S *s_x /*[2]*/) {
    if (threadIdx.x == 0) s_x[0] = total;
    cl.sync();
    if (threadIdx.x == 0) {
        S c = S::identity();
        const unsigned me = cl.block_rank();
        for (unsigned k = 0; k < me; ++k) c = S::combine(c, *cl.map_shared_rank(&s_x[0], k));
        s_x[1] = c;
    }
    __syncthreads();
    return s_x[1];
}

// grid = 2 clusters: cluster 0 scans the tiles in reverse (S, R), cluster 1 forward (L)
__global__ void __cluster_dims__(SC_CLUSTER, 1, 1) __launch_bounds__(SC_THREADS)
k_tile_scan1(const TileW *__restrict__ tw, int64_t n_tiles_max, const int64_t *__restrict__ tile_base, int n_seg,
             TileC1 *__restrict__ c1, ShardRec1 *__restrict__ rec /* shards: the folded sequence, else nullptr */) {
    __shared__ SegN<1> bufS[33], xS[2];
    __shared__ SegN<2> bufR[33], xR[2];
    __shared__ SegN<4> bufL[33], xL[2];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    const int64_t T = tile_base ? tile_base[n_seg] : n_tiles_max;
    const ScanRange rg = scan_range(T, rank);
    if (blockIdx.x < SC_CLUSTER) {   // reverse: S (restart at a cohort's last tile), chain of first fragments (E, m)
        SegN<1> carS = SegN<1>::identity();
        SegN<2> carR = SegN<2>::identity();
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
            for (int64_t r = rg.r0; r < rg.r1; ++r) {
                SegN<1> eS[SC_ITEMS];
                SegN<2> eR[SC_ITEMS];
                int last[SC_ITEMS];
                SegN<1> aggS = SegN<1>::identity();
                SegN<2> aggR = SegN<2>::identity();
#pragma unroll
                for (int k = 0; k < SC_ITEMS; ++k) {
                    const int64_t q = T - 1 - (r * SC_ROUND + (int64_t)threadIdx.x * SC_ITEMS + k);
                    // branch-free: every field is loaded from a clamped record and masked afterwards, so that the loads of all
                    // SC_ITEMS records are in flight together
                    const bool in = q >= 0;
                    const TileW *x = tw + (in ? q : 0);
                    const double Wtot = x->Wtot, Ef = x->Ef;
                    const int mf = x->mf, rowsf = x->rowsf, nheads = x->nheads, fl = x->flags;
                    eS[k].v[0] = in ? Wtot : 0.0; eS[k].flag = (in && (fl & TF_LAST)) ? 1 : 0;
                    eR[k].v[0] = (in && rowsf) ? Ef : 0.0; eR[k].v[1] = (in && rowsf) ? (double)mf : 0.0; eR[k].flag = in && nheads > 0;
                    last[k] = in ? (fl & TF_LAST) : 0;
                    aggS = SegN<1>::combine(aggS, eS[k]); aggR = SegN<2>::combine(aggR, eR[k]);
                }
                SegN<1> stS = SegN<1>::combine(carS, round_prefix<1>(aggS, bufS));
                SegN<2> stR = SegN<2>::combine(carR, round_prefix<2>(aggR, bufR));
                if (phase) {
#pragma unroll
                    for (int k = 0; k < SC_ITEMS; ++k) {
                        const int64_t q = T - 1 - (r * SC_ROUND + (int64_t)threadIdx.x * SC_ITEMS + k);
                        if (q >= 0) {
                            TileC1 *o = c1 + q;
                            o->S = last[k] ? 0.0 : stS.v[0];
                            o->RE = last[k] ? 0.0 : stR.v[0];
                            o->Rm = last[k] ? 0 : (int)(stR.v[1] + 0.5);
                            if (rec != nullptr && !last[k] && !stR.flag) atomicOr(&o->cf, 1);   // the forward cluster owns bit 1 of the same word
                        }
                        stS = SegN<1>::combine(stS, eS[k]); stR = SegN<2>::combine(stR, eR[k]);
                    }
                }
                carS = SegN<1>::combine(carS, bufS[32]); carR = SegN<2>::combine(carR, bufR[32]);
            }
            if (phase == 0) { carS = cluster_carry(cl, carS, xS); carR = cluster_carry(cl, carR, xR); }
        }
        if (rec != nullptr && rank == SC_CLUSTER - 1 && threadIdx.x == 0) {
            rec->W = carS.v[0]; rec->RE = carR.v[0]; rec->Rm = carR.v[1]; rec->Rflag = carR.flag; rec->pad0 = 0;
        }
    } else {                 // forward: chain of last fragments (W, E, m, rows)
        SegN<4> carL = SegN<4>::identity();
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
            for (int64_t r = rg.r0; r < rg.r1; ++r) {
                SegN<4> e[SC_ITEMS];
                int open[SC_ITEMS];
                SegN<4> agg = SegN<4>::identity();
#pragma unroll
                for (int k = 0; k < SC_ITEMS; ++k) {
                    const int64_t q = r * SC_ROUND + (int64_t)threadIdx.x * SC_ITEMS + k;
                    const bool in = q < T;
                    const TileW *x = tw + (in ? q : 0);
                    const double Wl = x->Wl, El = x->El;
                    const int ml = x->ml, rowsl = x->rowsl, nheads = x->nheads, rowsf = x->rowsf, fl = x->flags;
                    e[k].v[0] = in ? Wl : 0.0; e[k].v[1] = in ? El : 0.0; e[k].v[2] = in ? (double)ml : 0.0; e[k].v[3] = in ? (double)rowsl : 0.0;
                    e[k].flag = in && nheads > 0;
                    open[k] = in && rowsf > 0 && !(fl & TF_FIRST);
                    agg = SegN<4>::combine(agg, e[k]);
                }
                SegN<4> st = SegN<4>::combine(carL, round_prefix<4>(agg, bufL));
                if (phase) {
#pragma unroll
                    for (int k = 0; k < SC_ITEMS; ++k) {
                        const int64_t q = r * SC_ROUND + (int64_t)threadIdx.x * SC_ITEMS + k;
                        if (q < T) {
                            TileC1 *o = c1 + q;
                            o->LW = open[k] ? st.v[0] : 0.0; o->LE = open[k] ? st.v[1] : 0.0;
                            o->Lm = open[k] ? (int)(st.v[2] + 0.5) : 0; o->Lrows = open[k] ? (int)(st.v[3] + 0.5) : 0;
                            if (rec != nullptr && open[k] && !st.flag) atomicOr(&o->cf, 2);
                        }
                        st = SegN<4>::combine(st, e[k]);
                    }
                }
                carL = SegN<4>::combine(carL, bufL[32]);
            }
            if (phase == 0) carL = cluster_carry(cl, carL, xL);
        }
        if (rec != nullptr && rank == SC_CLUSTER - 1 && threadIdx.x == 0) {
            rec->LW = carL.v[0]; rec->LE = carL.v[1]; rec->Lm = carL.v[2]; rec->Lrows = carL.v[3]; rec->Lflag = carL.flag; rec->pad1 = 0;
        }
    }
    cl.sync();   // no CTA leaves while a later one may still read its total
}

// shards: a tile's shard-local scan values plus what the other shards contribute (k_shard_ctx1 / k_shard_finish)
__device__ __forceinline__ void shard_fix(TileC1 &c, const ShardCtx *__restrict__ ctx) {
    if (ctx == nullptr) return;
    c.S += ctx->carS;
    if (c.cf & 1) { c.RE += ctx->carRE; c.Rm += ctx->carRm; }
    if (c.cf & 2) { c.LW += ctx->carLW; c.LE += ctx->carLE; c.Lm += ctx->carLm; c.Lrows += ctx->carLrows; }
}
__device__ __forceinline__ void shard_fix(TileC2 &c, const ShardCtx *__restrict__ ctx) {
    if (ctx == nullptr) return;
    c.C += ctx->carC;
    if (c.cr) { c.AR += ctx->carAR; c.FR += ctx->carFR; }
    if (c.cfw) c.FL += ctx->carFL;
}

// ---- per-row terms of a tile (k_tile_terms): reverse scan -> every head publishes its group's
// (D, E, m) -> forward max scan (group starts) -> a_p, f_p of the event rows
static_assert(TS_TILE < 0xfff, "tile positions are packed into 12 bits");
struct TileTerms {
    double a[TS_ITEMS], f[TS_ITEMS];
    // packed (a register each would spill the sweep): bits 0-11 = 1 + start of the row's group inside the tile (0: the group
    // began in an earlier tile), bits 12-23 = last row of the row's group inside the tile (0xfff: it reaches past the tile)
    unsigned gt[TS_ITEMS];
    __device__ __forceinline__ int gs(int k) const { return (int)(gt[k] & 0xfffu) - 1; }
    __device__ __forceinline__ bool open(int k) const { return (gt[k] >> 12) == 0xfffu; }
    __device__ __forceinline__ int tpos(int k) const { return (int)(gt[k] >> 12); }
};
template <bool WITH_LOG>
__device__ __forceinline__ void tile_terms(const TileGeo &g, const TileRows &R, const TileW &tw, const TileC1 &c, int efron,
                                           double *s_D, double *s_E, int *s_m, RevT *s_rev, MaxT *s_max, TileTerms &X,
                                           double &sum_log) {
    const int t = threadIdx.x;
    {   // reverse scan; the tables take the place of the staged keys / weights (block_prefix synchronises first)
        RevT agg = RevT::identity();
#pragma unroll
        for (int k = TS_ITEMS - 1; k >= 0; --k) {
            RevT e;
            const double w = (double)R.w[k];
            e.W = w; e.E = (R.flg[k] & RF_EV) ? w : 0.0; e.m = (R.flg[k] & RF_EV) ? 1 : 0;
            e.tpos = (R.flg[k] & RF_TAIL) ? t * TS_ITEMS + k : INT_MAX;
            agg = RevT::combine(agg, e);
        }
        RevT acc = block_prefix<RevT, true>(agg, s_rev);
#pragma unroll
        for (int k = TS_ITEMS - 1; k >= 0; --k) {
            const int j = t * TS_ITEMS + k;
            RevT e;
            const double w = (double)R.w[k];
            e.W = w; e.E = (R.flg[k] & RF_EV) ? w : 0.0; e.m = (R.flg[k] & RF_EV) ? 1 : 0;
            e.tpos = (R.flg[k] & RF_TAIL) ? j : INT_MAX;
            acc = RevT::combine(acc, e);
            X.gt[k] = (acc.tpos == INT_MAX ? 0xfffu : (unsigned)acc.tpos) << 12;
            if (R.flg[k] & RF_HEAD) {   // a group that reaches the tile's end continues in later tiles (R)
                const bool open = acc.tpos == INT_MAX;
                s_D[sk(j)] = c.S + acc.W;
                s_E[sk(j)] = acc.E + (open ? c.RE : 0.0);
                s_m[sk4(j)] = acc.m + (open ? c.Rm : 0);
            }
        }
    }
    {   // group starts
        MaxT agg = MaxT::identity();
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) if (R.flg[k] & RF_HEAD) agg.h = t * TS_ITEMS + k;
        MaxT acc = block_prefix<MaxT, false>(agg, s_max);   // its barriers also order the table writes before the reads
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k) {
            if (R.flg[k] & RF_HEAD) acc.h = t * TS_ITEMS + k;
            X.gt[k] |= (unsigned)(acc.h + 1);
        }
    }
    // the group of the rows before the first head began in an earlier tile (L); it may also reach past this tile (R)
    const double Df = c.S + tw.Wtot + c.LW, Ef = c.LE + tw.Ef + (tw.nheads ? 0.0 : c.RE);
    const int mf = c.Lm + tw.mf + (tw.nheads ? 0 : c.Rm);
    sum_log = 0.0;
    // the logs of a thread's denominators as ONE logarithm: exponents summed as integers, mantissas (each in [1, 2)) multiplied
    double lprod = 1.0;
    int lexp = 0;
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        X.a[k] = 0.0; X.f[k] = 0.0;
        if (R.flg[k] & RF_EV) {
            const int j = t * TS_ITEMS + k, h = X.gs(k);
            const double D = h >= 0 ? s_D[sk(h)] : Df, E = h >= 0 ? s_E[sk(h)] : Ef;
            // a group's events come first on every shard, so the events of the group before this row are j - h of this tile
            // or, when the group began earlier, all the events of the earlier tiles (on one GPU that equals c.Lrows)
            const int m = h >= 0 ? s_m[sk4(h)] : mf, l = h >= 0 ? j - h : c.Lm + j;
            double den = D, frac = 0.0;
            if (efron && l > 0) { frac = (double)l * fast_rcp((double)m); den -= frac * E; }
            const double a = fast_rcp(den);
            X.a[k] = a; X.f[k] = frac * a;
            if (WITH_LOG) {
                const int hi = __double2hiint(den);
                lexp += ((hi >> 20) & 0x7ff) - 1023;
                lprod *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(den));
            }
        }
    }
    if (WITH_LOG) sum_log = (double)lexp * 0.6931471805599453 + fast_log(lprod);   // lprod in [1, 2^TS_ITEMS)
}

// Besides the tile's sums the sweep leaves, per row, everything of the gradient that does not depend on the SECOND tile scan:
//   g0 = d - w (P_tile - d F_tile)   (P_tile, F_tile: the in-tile prefix of a / group sum of f at the row's group end)
//   rfl = bit 0: event row, bit 1: the row's group reaches past the tile, bit 2: the group began in an earlier tile
// so that the last sweep (k_tile_apply) is element-wise: grad = scale (g0 - w (C + [bit 1] AR) + w d ([bit 1] FR + [bit 2] FL)).
// Earlier versions recomputed the scans and the terms in the gradient sweep (k_tile_grad: 320 us at 16.7M rows against
// 180 us for this kernel, 80 registers with the scatter's latency on top).
__global__ void __launch_bounds__(TS_THREADS, 3)
k_tile_terms(const uint32_t *__restrict__ keys_s, const float *__restrict__ w, const int64_t *__restrict__ seg_off,
             const int64_t *__restrict__ tile_base, int n_seg, int64_t n, const TileW *__restrict__ tws,
             const TileC1 *__restrict__ c1, int efron, const ShardCtx *__restrict__ ctx, TileA *__restrict__ ta,
             float *__restrict__ g0, uint8_t *__restrict__ rfl) {
    // (keys, weights) are staged only until the rows sit in registers; the group tables then take their place
    __shared__ double s_pool[2 * TS_DN + (TS_KN + 1) / 2];
    double *s_D = s_pool, *s_E = s_pool + TS_DN;
    int *s_m = reinterpret_cast<int *>(s_pool + 2 * TS_DN);
    uint32_t *s_key = reinterpret_cast<uint32_t *>(s_pool);
    float *s_w = reinterpret_cast<float *>(s_key + TS_KN);
    __shared__ RevT s_rev[TS_NW];
    __shared__ MaxT s_max[TS_NW];
    __shared__ FwdT s_fwd[TS_NW];
    __shared__ double s_red[6][TS_NW];
    __shared__ int s_redi[2][TS_NW];
    const TileGeo g = tile_geo(blockIdx.x, seg_off, tile_base, n_seg, n, ctx);
    if (!g.valid) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const TileW tw = tws[blockIdx.x];
    TileC1 c = c1[blockIdx.x];
    shard_fix(c, ctx);
    TileRows R;
    tile_load(g, keys_s, w, s_key, s_w, R);
    TileTerms X;
    double sl;
    tile_terms<true>(g, R, tw, c, efron, s_D, s_E, s_m, s_rev, s_max, X, sl);
    // reductions: the log-denominators and the event / event-time counts.  The sums of a, f over the tile and over its two
    // fragments are read off the forward scan below (they are prefixes / group-prefixes at the fragment ends).
    int nt = 0, ne = 0;
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t * TS_ITEMS + k;
        // a distinct event time is counted at its event with l = 0: the group's head, or (shards only: the head is a censored
        // row of the previous shard) the first row of a tile that continues a group without events so far
        ne += (R.flg[k] & RF_EV) ? 1 : 0;
        nt += ((R.flg[k] & (RF_EV | RF_HEAD)) == (RF_EV | RF_HEAD) || ((R.flg[k] & RF_EV) && j == 0 && X.gs(k) < 0 && c.Lm == 0)) ? 1 : 0;
    }
    sl = warp_sum(sl);
    nt = (int)warp_sum((long long)nt); ne = (int)warp_sum((long long)ne);
    if (lane == 0) { s_red[0][warp] = sl; s_redi[0][warp] = nt; s_redi[1][warp] = ne; }
    // forward: P (prefix of a inside the tile), group-prefixes of a and f; (P, F) at every row take the place of the group tables
    FwdT agg = FwdT::identity();
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        FwdT e; e.A = X.a[k]; e.As = X.a[k]; e.F = X.f[k]; e.head = (R.flg[k] & RF_HEAD) ? 1 : 0;
        agg = FwdT::combine(agg, e);
    }
    FwdT run = block_prefix<FwdT, false>(agg, s_fwd);   // its barriers: every thread is done with the (D, E, m) tables
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        FwdT e; e.A = X.a[k]; e.As = X.a[k]; e.F = X.f[k]; e.head = (R.flg[k] & RF_HEAD) ? 1 : 0;
        run = FwdT::combine(run, e);
        const int j = t * TS_ITEMS + k;
        if (j < g.rows) { s_D[sk(j)] = run.A; s_E[sk(j)] = run.F; }
    }
    if (t == (g.rows - 1) / TS_ITEMS) s_red[1][0] = run.As;   // sum of a from the tile's last head on (rows past the end are neutral)
    __syncthreads();
    if (t == 0) {
        double sumL = 0.0;
        int rt = 0, re = 0;
        for (int q = 0; q < TS_NW; ++q) { sumL += s_red[0][q]; rt += s_redi[0][q]; re += s_redi[1][q]; }
        TileA o;
        o.sumA = s_D[sk(g.rows - 1)]; o.sumL = sumL;
        o.Af = tw.rowsf > 0 ? s_D[sk(tw.rowsf - 1)] : 0.0; o.Ff = tw.rowsf > 0 ? s_E[sk(tw.rowsf - 1)] : 0.0;
        o.Al = s_red[1][0]; o.Fl = s_E[sk(g.rows - 1)];
        o.n_times = rt; o.n_ev = re; o.pad0 = 0; o.pad1 = 0;
        ta[blockIdx.x] = o;
    }
    float gv[TS_ITEMS];
    unsigned fv[TS_ITEMS];
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t * TS_ITEMS + k;
        gv[k] = 0.f; fv[k] = 0u;
        if (j < g.rows) {
            const bool open = X.open(k);                        // the row's group reaches past the tile
            const int e = open ? g.rows - 1 : X.tpos(k);        // its last row inside the tile
            const double d = (R.flg[k] & RF_EV) ? 1.0 : 0.0;
            gv[k] = (float)(d - (double)R.w[k] * (s_D[sk(e)] - d * s_E[sk(e)]));
            fv[k] = ((R.flg[k] & RF_EV) ? 1u : 0u) | (open ? 2u : 0u) | (X.gs(k) < 0 ? 4u : 0u);
        }
    }
    const int64_t pb = g.p0 + (int64_t)t * TS_ITEMS;
    if ((g.p0 & 7) == 0 && t * TS_ITEMS + TS_ITEMS <= g.rows) {   // whole 32-byte / 8-byte runs (always, for one cohort)
        *reinterpret_cast<float4 *>(g0 + pb) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        *reinterpret_cast<float4 *>(g0 + pb + 4) = make_float4(gv[4], gv[5], gv[6], gv[7]);
        *reinterpret_cast<uint2 *>(rfl + pb) = make_uint2(fv[0] | (fv[1] << 8) | (fv[2] << 16) | (fv[3] << 24),
                                                           fv[4] | (fv[5] << 8) | (fv[6] << 16) | (fv[7] << 24));
    } else {
#pragma unroll
        for (int k = 0; k < TS_ITEMS; ++k)
            if (t * TS_ITEMS + k < g.rows) { g0[pb + k] = gv[k]; rfl[pb + k] = (uint8_t)fv[k]; }
    }
}

// second tile scan + per-cohort loss / scale / header.  grid = 2 clusters: cluster 0 reverse (AR, FR), cluster 1 forward
// (C, FL, loss).
__global__ void __cluster_dims__(SC_CLUSTER, 1, 1) __launch_bounds__(SC_THREADS)
k_tile_scan2(const TileW *__restrict__ tw, const TileA *__restrict__ ta, int64_t n_tiles_max, const int64_t *__restrict__ tile_base,
             const int64_t *__restrict__ seg_off, int n_seg, int64_t n, int ties, int reduction, SegAcc *acc,
             TileC2 *__restrict__ c2, float *__restrict__ out_loss, b200surv_cox_header *__restrict__ hdrs,
             ShardRec2 *__restrict__ rec /* shards: the folded sequence (loss and header come from k_shard_finish), else nullptr */) {
    __shared__ SegN<2> buf[33], xA[2];
    __shared__ SegN<4> bufC[33], xC[2];
    __shared__ SegN<1> bufF[33], xF[2];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    const int64_t T = tile_base ? tile_base[n_seg] : n_tiles_max;
    const ScanRange rg = scan_range(T, rank);
    const int t = threadIdx.x;
    if (blockIdx.x < SC_CLUSTER) {   // reverse: chain of first fragments (A, F)
        SegN<2> car = SegN<2>::identity();
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
            for (int64_t r = rg.r0; r < rg.r1; ++r) {
                SegN<2> e[SC_ITEMS];
                int last[SC_ITEMS];
                SegN<2> agg = SegN<2>::identity();
#pragma unroll
                for (int k = 0; k < SC_ITEMS; ++k) {
                    const int64_t q = T - 1 - (r * SC_ROUND + (int64_t)t * SC_ITEMS + k);
                    const bool in = q >= 0;
                    const TileW *x = tw + (in ? q : 0); const TileA *y = ta + (in ? q : 0);
                    const int rowsf = x->rowsf, nheads = x->nheads, fl = x->flags;
                    const double Af = y->Af, Ff = y->Ff;
                    e[k].v[0] = (in && rowsf) ? Af : 0.0; e[k].v[1] = (in && rowsf) ? Ff : 0.0; e[k].flag = in && nheads > 0;
                    last[k] = in ? (fl & TF_LAST) : 0;
                    agg = SegN<2>::combine(agg, e[k]);
                }
                SegN<2> st = SegN<2>::combine(car, round_prefix<2>(agg, buf));
                if (phase) {
#pragma unroll
                    for (int k = 0; k < SC_ITEMS; ++k) {
                        const int64_t q = T - 1 - (r * SC_ROUND + (int64_t)t * SC_ITEMS + k);
                        if (q >= 0) {
                            c2[q].AR = last[k] ? 0.0 : st.v[0]; c2[q].FR = last[k] ? 0.0 : st.v[1];
                            c2[q].cr = (!last[k] && !st.flag) ? 1 : 0;
                        }
                        st = SegN<2>::combine(st, e[k]);
                    }
                }
                car = SegN<2>::combine(car, buf[32]);
            }
            if (phase == 0) car = cluster_carry(cl, car, xA);
        }
        if (rec != nullptr && rank == SC_CLUSTER - 1 && t == 0) { rec->AR = car.v[0]; rec->FR = car.v[1]; rec->ARflag = car.flag; rec->pad0 = 0; }
        cl.sync();
        return;
    }
    {   // forward: C and the cohort's loss sums (restart at a cohort's first tile); chain of last fragments (F)
        SegN<4> carC = SegN<4>::identity();
        SegN<1> carF = SegN<1>::identity();
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
            for (int64_t r = rg.r0; r < rg.r1; ++r) {
                SegN<4> e[SC_ITEMS];
                SegN<1> ef[SC_ITEMS];
                int fl[SC_ITEMS];   // bit 0: first tile of a cohort, 1: last, 2: the first row continues an earlier tile's group
                SegN<4> aggC = SegN<4>::identity();
                SegN<1> aggF = SegN<1>::identity();
#pragma unroll
                for (int k = 0; k < SC_ITEMS; ++k) {
                    const int64_t q = r * SC_ROUND + (int64_t)t * SC_ITEMS + k;
                    const bool in = q < T;
                    const TileW *x = tw + (in ? q : 0); const TileA *y = ta + (in ? q : 0);
                    const int xf = x->flags, nheads = x->nheads, rowsf = x->rowsf, ynt = y->n_times, yne = y->n_ev;
                    const double sumA = y->sumA, sumL = y->sumL, Fl = y->Fl;
                    e[k].v[0] = in ? sumA : 0.0; e[k].v[1] = in ? sumL : 0.0; e[k].v[2] = in ? (double)ynt : 0.0; e[k].v[3] = in ? (double)yne : 0.0;
                    e[k].flag = (in && (xf & TF_FIRST)) ? 1 : 0;
                    ef[k].v[0] = in ? Fl : 0.0; ef[k].flag = in && nheads > 0;
                    fl[k] = in ? ((xf & (TF_FIRST | TF_LAST)) | ((rowsf > 0 && !(xf & TF_FIRST)) ? 4 : 0)) : 0;
                    aggC = SegN<4>::combine(aggC, e[k]); aggF = SegN<1>::combine(aggF, ef[k]);
                }
                SegN<4> stC = SegN<4>::combine(carC, round_prefix<4>(aggC, bufC));
                SegN<1> stF = SegN<1>::combine(carF, round_prefix<1>(aggF, bufF));
                if (phase) {
#pragma unroll
                    for (int k = 0; k < SC_ITEMS; ++k) {
                        const int64_t q = r * SC_ROUND + (int64_t)t * SC_ITEMS + k;
                        if (q < T) {
                            c2[q].C = (fl[k] & TF_FIRST) ? 0.0 : stC.v[0];
                            c2[q].FL = (fl[k] & 4) ? stF.v[0] : 0.0;
                            c2[q].cfw = ((fl[k] & 4) && !stF.flag) ? 1 : 0;
                        }
                        stC = SegN<4>::combine(stC, e[k]); stF = SegN<1>::combine(stF, ef[k]);
                        if (rec == nullptr && q < T && (fl[k] & TF_LAST)) {   // the cohort's totals are complete: its sums of log-denominators / event times
                            const TileGeo g = tile_geo(q, seg_off, tile_base, n_seg, n);
                            acc[g.seg].sum_log = stC.v[1];
                            acc[g.seg].n_times = (unsigned long long)(stC.v[2] + 0.5);
                        }
                    }
                }
                carC = SegN<4>::combine(carC, bufC[32]); carF = SegN<1>::combine(carF, bufF[32]);
            }
            if (phase == 0) { carC = cluster_carry(cl, carC, xC); carF = cluster_carry(cl, carF, xF); }
        }
        if (rec != nullptr) {
            if (rank == SC_CLUSTER - 1 && t == 0) {
                rec->A = carC.v[0]; rec->sum_log = carC.v[1]; rec->n_times = (long long)(carC.v[2] + 0.5);
                rec->n_ev = (long long)acc->n_ev; rec->sum_eta = acc->sum_eta;
                rec->FL = carF.v[0]; rec->FLflag = carF.flag; rec->flags = acc->flags;
            }
            cl.sync();
            return;
        }
    }
    __threadfence();
    cl.sync();   // every cohort's sum_log / n_times is written (and no CTA leaves while its total may still be read)
    for (int s = t + (int)rank * SC_THREADS; s < n_seg; s += SC_THREADS * SC_CLUSTER) {
        SegAcc &A = acc[s];
        // sum_log holds the logs of the SHIFTED denominators: log sum w e^{shift} = log sum w + shift
        const double pll = A.sum_eta - (A.sum_log + (double)A.n_ev * (double)A.max_eta);
        const double n_ev = (double)A.n_ev, n_times = (double)A.n_times;
        double norm = 1.0;
        if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = n_ev;
        else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = (ties == B200SURV_TIES_EFRON) ? n_times : n_ev;
        double scale = A.n_ev > 0 ? -1.0 / norm : 0.0, loss = A.n_ev > 0 ? -pll / norm : 0.0;
        if (A.flags) { loss = __longlong_as_double(0x7ff8000000000000ll); scale = loss; }
        A.scale = scale;
        b200surv_cox_header *hdr = hdrs + s;
        hdr->flags = A.flags; hdr->mode = B200SURV_COX_SORTED; hdr->loss = (float)loss; hdr->scale = (float)scale;
        hdr->shift = A.max_eta; hdr->max_log_hz = A.max_eta; hdr->max_time = A.max_time; hdr->nbins = 0;
        hdr->n_events = (int64_t)A.n_ev; hdr->n_event_times = (int64_t)A.n_times; hdr->pll = pll; hdr->min_log_hz = 0.f;
        hdr->reserved = 0;
        out_loss[s] = (float)loss;
    }
}

// the state keeps the UNSCALED per-row gradient d loss / d log_hz for grad_out = 1 (cox_scale_grad multiplies by grad_out).
// Last sweep, element-wise: the tile's second-scan carries on top of what k_tile_terms left per row, scattered through the
// permutation (4-byte stores over n rows: marked evict-last so that they merge in L2; the streams are evict-first).
__global__ void __launch_bounds__(TS_THREADS, 6)
k_tile_apply(const float *__restrict__ g0, const uint8_t *__restrict__ rfl, const float *__restrict__ w,
             const uint32_t *__restrict__ idx_s, const int64_t *__restrict__ seg_off, const int64_t *__restrict__ tile_base, int n_seg,
             int64_t n, const TileC2 *__restrict__ c2, const SegAcc *__restrict__ acc, const ShardCtx *__restrict__ ctx,
             float *__restrict__ grad_unit) {
    const TileGeo g = tile_geo(blockIdx.x, seg_off, tile_base, n_seg, n, ctx);
    if (!g.valid) return;
    TileC2 cc = c2[blockIdx.x];
    shard_fix(cc, ctx);
    const double scale = acc[g.seg].scale;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int t = threadIdx.x;
    float gv[TS_ITEMS], wv[TS_ITEMS];
    uint32_t ri[TS_ITEMS];
    unsigned fv[TS_ITEMS];
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t + k * TS_THREADS;
        const int64_t p = g.p0 + (j < g.rows ? j : 0);
        ri[k] = ldg_hint_u32(idx_s + p, pol_stream);
        gv[k] = ldg_hint_f32(g0 + p, pol_stream);
        wv[k] = ldg_hint_f32(w + p, pol_stream);
        fv[k] = rfl[p];
    }
#pragma unroll
    for (int k = 0; k < TS_ITEMS; ++k) {
        const int j = t + k * TS_THREADS;
        if (j < g.rows) {
            const double wk = (double)wv[k];
            double gr = (double)gv[k] - wk * (cc.C + ((fv[k] & 2u) ? cc.AR : 0.0));
            if (fv[k] & 1u) gr += wk * (((fv[k] & 2u) ? cc.FR : 0.0) + ((fv[k] & 4u) ? cc.FL : 0.0));
            stg_hint_f32(grad_unit + ri[k], (float)(scale * gr), pol_keep);
        }
    }
}

// ---- time-range shards: record / carry kernels (one thread each: <= 64 shards)
__global__ void k_shard_rec0(const SegAcc *__restrict__ acc, int64_t n, ShardRec0 *__restrict__ rec) {
    if (threadIdx.x || blockIdx.x) return;
    ShardRec0 r;
    r.max_eta = acc->max_eta; r.min_time = -acc->neg_min_time; r.max_time = acc->max_time; r.flags = acc->flags; r.n = n;
    *rec = r;
}
__device__ __forceinline__ const unsigned char *rec_at(const void *all, int r) {
    return static_cast<const unsigned char *>(all) + (size_t)r * SHARD_REC_BYTES;
}
// common exponent shift, the neighbours' edge times, the order check: shard r's times must not exceed shard r + 1's
__global__ void k_shard_ctx0(const void *__restrict__ all0, int rank, int world, SegAcc *acc, ShardCtx *ctx) {
    if (threadIdx.x || blockIdx.x) return;
    float mx = -INFINITY, mt = -INFINITY;
    unsigned flags = 0;
    for (int r = 0; r < world; ++r) {
        const ShardRec0 *x = reinterpret_cast<const ShardRec0 *>(rec_at(all0, r));
        mx = fmaxf(mx, x->max_eta); mt = fmaxf(mt, x->max_time); flags |= x->flags;
        if (r + 1 < world) {
            const ShardRec0 *y = reinterpret_cast<const ShardRec0 *>(rec_at(all0, r + 1));
            if (!(x->max_time <= y->min_time)) flags |= B200SURV_COXF_NOT_PARTITIONED;
        }
    }
    acc->max_eta = mx; acc->max_time = mt; acc->flags = flags;
    ShardCtx c;
    memset(&c, 0, sizeof(c));
    c.has_prev = rank > 0; c.has_next = rank + 1 < world;
    if (c.has_prev) c.key_prev = __float_as_uint(reinterpret_cast<const ShardRec0 *>(rec_at(all0, rank - 1))->max_time + 0.f) << 1;
    if (c.has_next) c.key_next = __float_as_uint(reinterpret_cast<const ShardRec0 *>(rec_at(all0, rank + 1))->min_time + 0.f) << 1;
    *ctx = c;
}
// carries of the first tile scan: the later shards' weight and chain of first fragments, the earlier shards' chain of last
// fragments, folded in scan order with the restart rule of SegN::combine
__global__ void k_shard_ctx1(const void *__restrict__ all1, int rank, int world, ShardCtx *ctx) {
    if (threadIdx.x || blockIdx.x) return;
    double S = 0.0, RE = 0.0, Rm = 0.0;
    for (int r = world - 1; r > rank; --r) {
        const ShardRec1 *x = reinterpret_cast<const ShardRec1 *>(rec_at(all1, r));
        S += x->W;
        if (x->Rflag) { RE = x->RE; Rm = x->Rm; } else { RE += x->RE; Rm += x->Rm; }
    }
    double LW = 0.0, LE = 0.0, Lm = 0.0, Lrows = 0.0;
    for (int r = 0; r < rank; ++r) {
        const ShardRec1 *x = reinterpret_cast<const ShardRec1 *>(rec_at(all1, r));
        if (x->Lflag) { LW = x->LW; LE = x->LE; Lm = x->Lm; Lrows = x->Lrows; }
        else { LW += x->LW; LE += x->LE; Lm += x->Lm; Lrows += x->Lrows; }
    }
    ctx->carS = S; ctx->carRE = RE; ctx->carRm = (int)(Rm + 0.5);
    ctx->carLW = LW; ctx->carLE = LE; ctx->carLm = (int)(Lm + 0.5); ctx->carLrows = (int)(Lrows + 0.5);
}
// carries of the second tile scan, and the cohort's loss / scale / header from the shards' sums (every shard computes the
// same values in the same order)
__global__ void k_shard_finish(const void *__restrict__ all2, int rank, int world, int ties, int reduction, SegAcc *acc,
                               ShardCtx *ctx, float *__restrict__ out_loss, b200surv_cox_header *__restrict__ hdr) {
    if (threadIdx.x || blockIdx.x) return;
    double AR = 0.0, FR = 0.0;
    for (int r = world - 1; r > rank; --r) {
        const ShardRec2 *x = reinterpret_cast<const ShardRec2 *>(rec_at(all2, r));
        if (x->ARflag) { AR = x->AR; FR = x->FR; } else { AR += x->AR; FR += x->FR; }
    }
    double C = 0.0, FL = 0.0;
    for (int r = 0; r < rank; ++r) {
        const ShardRec2 *x = reinterpret_cast<const ShardRec2 *>(rec_at(all2, r));
        C += x->A;
        if (x->FLflag) FL = x->FL; else FL += x->FL;
    }
    ctx->carAR = AR; ctx->carFR = FR; ctx->carC = C; ctx->carFL = FL;
    double sum_eta = 0.0, sum_log = 0.0;
    long long n_ev = 0, n_times = 0;
    unsigned flags = acc->flags;
    for (int r = 0; r < world; ++r) {
        const ShardRec2 *x = reinterpret_cast<const ShardRec2 *>(rec_at(all2, r));
        sum_eta += x->sum_eta; sum_log += x->sum_log; n_ev += x->n_ev; n_times += x->n_times; flags |= x->flags;
    }
    const double pll = sum_eta - (sum_log + (double)n_ev * (double)acc->max_eta);
    double norm = 1.0;
    if (reduction == B200SURV_REDUCE_MEAN_EVENTS) norm = (double)n_ev;
    else if (reduction == B200SURV_REDUCE_MEAN_TERMS) norm = (ties == B200SURV_TIES_EFRON) ? (double)n_times : (double)n_ev;
    double scale = n_ev > 0 ? -1.0 / norm : 0.0, loss = n_ev > 0 ? -pll / norm : 0.0;
    if (flags) { loss = __longlong_as_double(0x7ff8000000000000ll); scale = loss; }
    acc->scale = scale; acc->flags = flags; acc->sum_eta = sum_eta; acc->sum_log = sum_log;
    acc->n_ev = (unsigned long long)n_ev; acc->n_times = (unsigned long long)n_times;
    hdr->flags = flags; hdr->mode = B200SURV_COX_SORTED; hdr->loss = (float)loss; hdr->scale = (float)scale;
    hdr->shift = acc->max_eta; hdr->max_log_hz = acc->max_eta; hdr->max_time = acc->max_time; hdr->nbins = 0;
    hdr->n_events = n_ev; hdr->n_event_times = n_times; hdr->pll = pll; hdr->min_log_hz = 0.f; hdr->reserved = 0;
    out_loss[0] = (float)loss;
}

struct SortedLayout {
    size_t off_acc, off_ctx, off_keys, off_vals, off_keys_s, off_idx_s, off_segid, off_w, off_tbase, off_tw, off_c1, off_ta, off_c2, off_tmp, total;
    int64_t tiles_max;
};

SortedLayout sorted_layout(int64_t n, int64_t n_seg) {
    SortedLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.tiles_max = (int64_t)((N + TS_TILE - 1) / TS_TILE) + (n_seg > 1 ? n_seg : 0);   // every cohort starts a new tile
    const size_t T = (size_t)L.tiles_max;
    L.off_acc = take((size_t)n_seg * sizeof(SegAcc));
    L.off_ctx = take(sizeof(ShardCtx));
    L.off_keys = take(N * 4 + 4); L.off_vals = take(N * 4); L.off_keys_s = take(N * 4 + 4); L.off_idx_s = take(N * 4);
    L.off_segid = take(n_seg > 1 ? N * 4 : 4);
    L.off_w = take(N * 4);
    L.off_tbase = take((size_t)(n_seg + 1) * 8);
    L.off_tw = take(T * sizeof(TileW)); L.off_c1 = take(T * sizeof(TileC1)); L.off_ta = take(T * sizeof(TileA));
    L.off_c2 = take(T * sizeof(TileC2));
    L.off_tmp = take(sortscan::radix_sort_temp_bytes((int64_t)N));
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
    int64_t *tbase = seg_off ? reinterpret_cast<int64_t *>(w8 + L.off_tbase) : nullptr;
    TileW *tw = reinterpret_cast<TileW *>(w8 + L.off_tw);
    TileC1 *c1 = reinterpret_cast<TileC1 *>(w8 + L.off_c1);
    TileA *ta = reinterpret_cast<TileA *>(w8 + L.off_ta);
    TileC2 *c2 = reinterpret_cast<TileC2 *>(w8 + L.off_c2);
    void *tmp = w8 + L.off_tmp;
    int grid = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    if (grid > cap) grid = cap;
    const int nseg = (int)n_seg;
    const unsigned tiles = (unsigned)L.tiles_max;
    const int efron = ties == B200SURV_TIES_EFRON ? 1 : 0;

    b200surv_cox_header *hdrs = static_cast<b200surv_cox_header *>(state);
    float *grad_unit = reinterpret_cast<float *>(hdrs + n_seg);
    int32_t rc;

    // B200SURV_SORTED_TRACE=1: per-phase device times of this call on stderr (events on the caller's stream; debugging aid)
    static const bool trace = getenv("B200SURV_SORTED_TRACE") != nullptr;
    cudaEvent_t ev[10];
    int nev = 0;
    auto mark = [&]() { if (trace && nev < 10) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], st); ++nev; } };
    mark();
    k_init_acc<<<(nseg + 255) / 256, 256, 0, st>>>(acc, nseg);
    // keys are generated into (keys_s, idx_s); radix_sort_pairs2 reports which buffer pair holds the result
    unsigned *hist0 = nullptr;   // the key kernel counts the sort's first digit while it has the keys in registers
    rc = sortscan::radix_sort_prepare(n, tmp, st, &hist0);
    if (rc) return rc;
    k_make_keys<<<grid, 256, 0, st>>>(log_hz, time, event, seg_off, nseg, n, keys_s, idx_s, segid, acc, hist0);
    mark();
    const int seg_bits = n_seg == 1 ? 0 : (n_seg <= 256 ? 8 : 16);
    int in_first = 1;
    rc = sortscan::radix_sort_pairs2(keys_s, idx_s, keys, vals, n, 32, n_seg > 1 ? segid : nullptr, seg_bits, tmp, st, &in_first, true,
                                     true);
    if (rc) return rc;
    mark();
    const uint32_t *ks = in_first ? keys_s : keys, *is = in_first ? idx_s : vals;
    mark();
    if (seg_off) k_tile_base<<<1, 1024, 0, st>>>(seg_off, nseg, tbase);
    k_tile_w<<<tiles, TS_THREADS, 0, st>>>(log_hz, ks, is, wv, seg_off, tbase, nseg, n, nullptr, acc, tw);
    mark();
    k_tile_scan1<<<2 * SC_CLUSTER, SC_THREADS, 0, st>>>(tw, L.tiles_max, tbase, nseg, c1, nullptr);
    mark();
    // the buffer pair the sort did not end in is free: per-row partial gradient and flags between the last two sweeps
    float *g0 = reinterpret_cast<float *>(in_first ? keys : keys_s);
    uint8_t *rfl = reinterpret_cast<uint8_t *>(in_first ? vals : idx_s);
    k_tile_terms<<<tiles, TS_THREADS, 0, st>>>(ks, wv, seg_off, tbase, nseg, n, tw, c1, efron, nullptr, ta, g0, rfl);
    mark();
    k_tile_scan2<<<2 * SC_CLUSTER, SC_THREADS, 0, st>>>(tw, ta, L.tiles_max, tbase, seg_off, nseg, n, ties, reduction, acc, c2, out_loss, hdrs, nullptr);
    mark();
    k_tile_apply<<<tiles, TS_THREADS, 0, st>>>(g0, rfl, wv, is, seg_off, tbase, nseg, n, c2, acc, nullptr, grad_unit);
    mark();
    if (trace) {
        static const char *names[] = {"keys", "sort", "weights", "tile_w", "scan1", "terms", "scan2", "grad"};
        cudaEventSynchronize(ev[nev - 1]);
        fprintf(stderr, "[sorted n=%lld]", (long long)n);
        for (int i = 0; i + 1 < nev; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            fprintf(stderr, " %s %.1f us", names[i], ms * 1e3f);
        }
        fprintf(stderr, "\n");
        for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    }
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(2 + (4 + seg_bits / 8) + (seg_off ? 1 : 0) + 5);   // acc, keys (+ first digit totals); one sweep per digit; tile kernels
    return B200SURV_OK;
}


// ---- time-range shards: the same kernels in four phases; the caller all-gathers one 128-byte record per shard between them
namespace {
struct ShardPtrs {
    SortedLayout L;
    SegAcc *acc; ShardCtx *ctx;
    uint32_t *keys, *vals, *keys_s, *idx_s;
    float *wv;
    TileW *tw; TileC1 *c1; TileA *ta; TileC2 *c2;
    void *tmp;
    int grid; unsigned tiles;
};
int32_t shard_ptrs(int64_t n, void *ws, size_t ws_bytes, ShardPtrs &P) {
    B200_REQUIRE(ws != nullptr, "workspace");
    B200_REQUIRE(n >= 1 && n < (int64_t)INT_MAX - 2, "every shard needs n in [1, 2^31) rows");
    P.L = sorted_layout(n, 1);
    if (ws_bytes < P.L.total) { set_error("cox sorted shard: workspace %zu < %zu", ws_bytes, P.L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    unsigned char *w8 = static_cast<unsigned char *>(ws);
    P.acc = reinterpret_cast<SegAcc *>(w8 + P.L.off_acc);
    P.ctx = reinterpret_cast<ShardCtx *>(w8 + P.L.off_ctx);
    P.keys = reinterpret_cast<uint32_t *>(w8 + P.L.off_keys); P.vals = reinterpret_cast<uint32_t *>(w8 + P.L.off_vals);
    P.keys_s = reinterpret_cast<uint32_t *>(w8 + P.L.off_keys_s); P.idx_s = reinterpret_cast<uint32_t *>(w8 + P.L.off_idx_s);
    P.wv = reinterpret_cast<float *>(w8 + P.L.off_w);
    P.tw = reinterpret_cast<TileW *>(w8 + P.L.off_tw); P.c1 = reinterpret_cast<TileC1 *>(w8 + P.L.off_c1);
    P.ta = reinterpret_cast<TileA *>(w8 + P.L.off_ta); P.c2 = reinterpret_cast<TileC2 *>(w8 + P.L.off_c2);
    P.tmp = w8 + P.L.off_tmp;
    P.grid = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    if (P.grid > cap) P.grid = cap;
    P.tiles = (unsigned)P.L.tiles_max;
    return B200SURV_OK;
}
// the radix sort ends in a buffer pair that depends only on the number of passes (32 key bits, no cohort passes): 4 passes
// of 8 bits end where they started
constexpr int SHARD_IN_FIRST = 1;
}  // namespace

int32_t cox_sorted_shard_keys(const float *log_hz, const float *time, const uint8_t *event, int64_t n, void *rec0_out, void *ws,
                              size_t ws_bytes, cudaStream_t st) {
    ShardPtrs P;
    int32_t rc = shard_ptrs(n, ws, ws_bytes, P);
    if (rc) return rc;
    k_init_acc<<<1, 256, 0, st>>>(P.acc, 1);
    unsigned *hist0 = nullptr;
    rc = sortscan::radix_sort_prepare(n, P.tmp, st, &hist0);
    if (rc) return rc;
    k_make_keys<<<P.grid, 256, 0, st>>>(log_hz, time, event, nullptr, 1, n, P.keys_s, P.idx_s, nullptr, P.acc, hist0);
    k_shard_rec0<<<1, 32, 0, st>>>(P.acc, n, static_cast<ShardRec0 *>(rec0_out));
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(3);
    return B200SURV_OK;
}

int32_t cox_sorted_shard_sort(int64_t n, void *ws, size_t ws_bytes, cudaStream_t st) {
    ShardPtrs P;
    int32_t rc = shard_ptrs(n, ws, ws_bytes, P);
    if (rc) return rc;
    int in_first = 1;
    rc = sortscan::radix_sort_pairs2(P.keys_s, P.idx_s, P.keys, P.vals, n, 32, nullptr, 0, P.tmp, st, &in_first, true, true);
    if (rc) return rc;
    if (in_first != SHARD_IN_FIRST) { set_error("cox sorted shard: unexpected sort buffer parity"); return B200SURV_UNSUPPORTED; }
    count_launches(4);   // one sweep per digit
    return B200SURV_OK;
}

int32_t cox_sorted_shard_reduce(const float *log_hz, int64_t n, const void *all_rec0, int rank, int world, void *rec1_out,
                                void *ws, size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(world >= 1 && world <= SHARD_MAX && rank >= 0 && rank < world, "rank / world");
    ShardPtrs P;
    int32_t rc = shard_ptrs(n, ws, ws_bytes, P);
    if (rc) return rc;
    k_shard_ctx0<<<1, 32, 0, st>>>(all_rec0, rank, world, P.acc, P.ctx);
    k_tile_w<<<P.tiles, TS_THREADS, 0, st>>>(log_hz, P.keys_s, P.idx_s, P.wv, nullptr, nullptr, 1, n, P.ctx, P.acc, P.tw);
    B200_CHECK_CUDA(cudaMemsetAsync(P.c1, 0, (size_t)P.L.tiles_max * sizeof(TileC1), st));   // the carry bits are OR-ed in
    k_tile_scan1<<<2 * SC_CLUSTER, SC_THREADS, 0, st>>>(P.tw, P.L.tiles_max, nullptr, 1, P.c1, static_cast<ShardRec1 *>(rec1_out));
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(3);
    return B200SURV_OK;
}

int32_t cox_sorted_shard_terms(int64_t n, int ties, const void *all_rec1, int rank, int world, void *rec2_out, void *ws,
                               size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(world >= 1 && world <= SHARD_MAX && rank >= 0 && rank < world, "rank / world");
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    ShardPtrs P;
    int32_t rc = shard_ptrs(n, ws, ws_bytes, P);
    if (rc) return rc;
    const int efron = ties == B200SURV_TIES_EFRON ? 1 : 0;
    k_shard_ctx1<<<1, 32, 0, st>>>(all_rec1, rank, world, P.ctx);
    k_tile_terms<<<P.tiles, TS_THREADS, 0, st>>>(P.keys_s, P.wv, nullptr, nullptr, 1, n, P.tw, P.c1, efron, P.ctx, P.ta,
                                                 reinterpret_cast<float *>(P.keys), reinterpret_cast<uint8_t *>(P.vals));
    k_tile_scan2<<<2 * SC_CLUSTER, SC_THREADS, 0, st>>>(P.tw, P.ta, P.L.tiles_max, nullptr, nullptr, 1, n, ties, 0, P.acc, P.c2, nullptr, nullptr,
                                           static_cast<ShardRec2 *>(rec2_out));
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(3);
    return B200SURV_OK;
}

int32_t cox_sorted_shard_finish(int64_t n, int ties, int reduction, const void *all_rec2, int rank, int world, float *out_loss,
                                void *state, size_t state_bytes, void *ws, size_t ws_bytes, cudaStream_t st) {
    B200_REQUIRE(world >= 1 && world <= SHARD_MAX && rank >= 0 && rank < world, "rank / world");
    B200_REQUIRE(ties == B200SURV_TIES_EFRON || ties == B200SURV_TIES_BRESLOW, "ties");
    B200_REQUIRE(reduction >= 0 && reduction <= 2, "reduction");
    ShardPtrs P;
    int32_t rc = shard_ptrs(n, ws, ws_bytes, P);
    if (rc) return rc;
    const size_t need = sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
    if (state_bytes < need) { set_error("cox sorted shard: state buffer %zu < %zu", state_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    b200surv_cox_header *hdr = static_cast<b200surv_cox_header *>(state);
    float *grad_unit = reinterpret_cast<float *>(hdr + 1);
    const int efron = ties == B200SURV_TIES_EFRON ? 1 : 0;
    k_shard_finish<<<1, 32, 0, st>>>(all_rec2, rank, world, ties, reduction, P.acc, P.ctx, out_loss, hdr);
    k_tile_apply<<<P.tiles, TS_THREADS, 0, st>>>(reinterpret_cast<const float *>(P.keys), reinterpret_cast<const uint8_t *>(P.vals), P.wv,
                                                 P.idx_s, nullptr, nullptr, 1, n, P.c2, P.acc, P.ctx, grad_unit);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(2);
    return B200SURV_OK;
}

}  // namespace b200surv
