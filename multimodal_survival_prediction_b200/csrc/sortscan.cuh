// Hand-written device-wide primitives for the SORTED Cox path and the C-index preprocessing (sm_100a):
//   * scan_lookback  -- single-pass inclusive scan with DECOUPLED LOOK-BACK over tiles of 2048 elements, forward or
//                       reverse, of a small tuple (two fp64 sums + one integer with add / min / max).  Fence-free:
//                       every 64-bit word a tile publishes (its aggregate, its inclusive prefix) is its own "valid"
//                       mark (SCAN_EMPTY until written), so there is no status flag to order the payload against
//                       (a gpu-scope fence costs microseconds on this part -- see DESIGN.md 3.1).  Tile ids are
//                       handed out by an atomic counter, so a tile's predecessors are always running or done.
//   * radix_sort_pairs -- stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass: per-tile digit
//                       histograms, one exclusive scan of the (digit, tile) matrix (scan_lookback), stable scatter
//                       with warp-level match ranking.
// They replace cub::DeviceScan / cub::DeviceRadixSort in cox_sorted.cu and cindex.cu.
#pragma once
#include <climits>

#include "common.cuh"

namespace b200surv {
namespace sortscan {

// ------------------------------------------------------------------------------------------------ scan
struct Tup {
    double a, b;   // summed
    long long i;   // combined with IOP
};
enum IntOp { I_ADD = 0, I_MIN = 1, I_MAX = 2 };

template <int IOP>
__device__ __forceinline__ Tup tup_identity() {
    Tup t;
    t.a = 0.0; t.b = 0.0;
    t.i = IOP == I_ADD ? 0ll : (IOP == I_MIN ? LLONG_MAX : LLONG_MIN);
    return t;
}
// `x` precedes `y` in scan order.  SEGA / SEGB turn the sum of a / b into a SEGMENTED sum that restarts at every
// segment head: an element marks a head through its integer (I_MAX: i >= 0, e.g. its own index, -1 otherwise;
// I_MIN: i != LLONG_MAX), so "y holds a head" is read off y.i and the operator stays associative
// ((f1,v1)+(f2,v2) = (f1|f2, f2 ? v2 : v1+v2)).  Sums of a tie group formed this way involve only the group's own
// terms -- differences of two global prefix sums cancel catastrophically when a group's weights are tiny.
template <int IOP, bool SEGA = false, bool SEGB = false>
__device__ __forceinline__ Tup tup_combine(const Tup &x, const Tup &y) {
    static_assert(!(SEGA || SEGB) || IOP != I_ADD, "segmented sums need head markers (I_MIN / I_MAX)");
    Tup t;
    const bool head = IOP == I_MAX ? y.i >= 0 : (IOP == I_MIN ? y.i != LLONG_MAX : false);
    t.a = (SEGA && head) ? y.a : x.a + y.a;
    t.b = (SEGB && head) ? y.b : x.b + y.b;
    t.i = IOP == I_ADD ? x.i + y.i : (IOP == I_MIN ? (x.i < y.i ? x.i : y.i) : (x.i > y.i ? x.i : y.i));
    return t;
}
__device__ __forceinline__ Tup tup_shfl_up(const Tup &v, int d) {
    Tup t;
    t.a = __shfl_up_sync(FULL, v.a, d); t.b = __shfl_up_sync(FULL, v.b, d); t.i = __shfl_up_sync(FULL, v.i, d);
    return t;
}
__device__ __forceinline__ Tup tup_shfl(const Tup &v, int src) {
    Tup t;
    t.a = __shfl_sync(FULL, v.a, src); t.b = __shfl_sync(FULL, v.b, src); t.i = __shfl_sync(FULL, v.i, src);
    return t;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048
constexpr unsigned long long SCAN_EMPTY = ~0ull;       // a NaN pattern / an integer no scan here produces
constexpr int SCAN_SPIN_MAX = 1 << 24;

// per launch: words[2][ntiles][3] (aggregate, inclusive) preset to SCAN_EMPTY, and a tile counter preset to 0
struct ScanState {
    unsigned long long *agg, *inc;
    unsigned *counter;
};
inline size_t scan_state_bytes(int64_t n) {
    const size_t ntiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    return align_up(2 * ntiles * 3 * sizeof(unsigned long long), 256) + 256;
}
inline ScanState scan_state_at(void *buf, int64_t n) {
    const size_t ntiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    ScanState s;
    s.agg = static_cast<unsigned long long *>(buf);
    s.inc = s.agg + ntiles * 3;
    s.counter = reinterpret_cast<unsigned *>(static_cast<unsigned char *>(buf) + align_up(2 * ntiles * 3 * sizeof(unsigned long long), 256));
    return s;
}
static __global__ void k_scan_state_init(ScanState s, size_t words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) s.agg[i] = SCAN_EMPTY;
    if (blockIdx.x == 0 && threadIdx.x == 0) *s.counter = 0;
}

__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void scan_st(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void scan_publish(unsigned long long *w, const Tup &t) {
    scan_st(w, (unsigned long long)__double_as_longlong(t.a));
    scan_st(w + 1, (unsigned long long)__double_as_longlong(t.b));
    scan_st(w + 2, (unsigned long long)t.i);
}
// all three words valid?  (each word is published on its own; a tuple is usable once none is SCAN_EMPTY)
__device__ __forceinline__ bool scan_try_read(const unsigned long long *w, Tup &t) {
    const unsigned long long x = scan_ld(w), y = scan_ld(w + 1), z = scan_ld(w + 2);
    t.a = __longlong_as_double((long long)x); t.b = __longlong_as_double((long long)y); t.i = (long long)z;
    return x != SCAN_EMPTY && y != SCAN_EMPTY && z != SCAN_EMPTY;
}

// Inclusive scan of load(p), p the PHYSICAL index; scan order = ascending p, or descending p when REVERSE.
// store(p, inclusive, element).  grid = number of tiles, SCAN_THREADS threads.
template <int IOP, bool REVERSE, bool SEGA, bool SEGB, typename Load, typename Store>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(int64_t n, Load load, Store store, ScanState st) {
    __shared__ Tup s_warp[SCAN_THREADS / 32];
    __shared__ Tup s_prefix;
    __shared__ unsigned s_tile;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(st.counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SCAN_TILE + (int64_t)t * SCAN_ITEMS;  // logical index of this thread's first item

    Tup item[SCAN_ITEMS];
    Tup run = tup_identity<IOP>();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t li = base + k;
        item[k] = li < n ? load(REVERSE ? n - 1 - li : li) : tup_identity<IOP>();
        run = tup_combine<IOP, SEGA, SEGB>(run, item[k]);
    }
    // block-wide exclusive scan of the thread aggregates
    Tup inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Tup u = tup_shfl_up(inc, d);
        if (lane >= d) inc = tup_combine<IOP, SEGA, SEGB>(u, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Tup wpre = tup_identity<IOP>(), tile_agg = tup_identity<IOP>();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) wpre = tup_combine<IOP, SEGA, SEGB>(wpre, s_warp[w]);
        tile_agg = tup_combine<IOP, SEGA, SEGB>(tile_agg, s_warp[w]);
    }
    Tup thread_excl = tup_shfl_up(inc, 1);
    if (lane == 0) thread_excl = tup_identity<IOP>();
    thread_excl = tup_combine<IOP, SEGA, SEGB>(wpre, thread_excl);

    // ---- decoupled look-back (warp 0): publish the aggregate, find the prefix of all earlier tiles
    if (warp == 0) {
        if (lane == 0) {
            if (tile == 0) scan_publish(st.inc, tile_agg);
            else scan_publish(st.agg + 3 * tile, tile_agg);
        }
        Tup prefix = tup_identity<IOP>();
        int64_t look = tile - 1;  // nearest tile not yet accounted for
        for (int guard = 0; look >= 0 && guard < SCAN_SPIN_MAX; ++guard) {
            // lane l examines tile look - l: 2 = inclusive prefix available, 1 = aggregate only, 0 = nothing yet
            const int64_t q = look - lane;
            Tup v = tup_identity<IOP>();
            int state = 2;  // lanes beyond tile 0 behave like "inclusive = identity"
            if (q >= 0) {
                if (scan_try_read(st.inc + 3 * q, v)) state = 2;
                else if (scan_try_read(st.agg + 3 * q, v)) state = 1;
                else state = 0;
            }
            // the usable run: lanes 0 .. first lane with an inclusive prefix, provided none before it is empty
            const unsigned m_inc = __ballot_sync(FULL, state == 2), m_none = __ballot_sync(FULL, state == 0);
            const int first_inc = m_inc ? __ffs(m_inc) - 1 : 32, first_none = m_none ? __ffs(m_none) - 1 : 32;
            const bool done = first_inc < first_none;                      // an inclusive prefix with no gap before it
            const int take = done ? first_inc + 1 : first_none;            // lanes [0, take) are combined (<= 32)
            // combine in scan order, farthest tile first: an ordered five-step tree (a serial chain of 32 shuffled
            // combines used to be the cost of every look-back hop)
            if (lane >= take) v = tup_identity<IOP>();
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Tup u = tup_shfl(v, (lane + d) & 31);
                if (lane + d < 32) v = tup_combine<IOP, SEGA, SEGB>(u, v);
            }
            prefix = tup_combine<IOP, SEGA, SEGB>(tup_shfl(v, 0), prefix);
            if (done) break;
            look -= take;  // go on behind the tiles taken (take == 0: poll the same, still empty tile again)
        }
        if (lane == 0) {
            if (tile != 0) scan_publish(st.inc + 3 * tile, tup_combine<IOP, SEGA, SEGB>(prefix, tile_agg));
            s_prefix = prefix;
        }
    }
    __syncthreads();
    Tup acc = tup_combine<IOP, SEGA, SEGB>(s_prefix, thread_excl);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t li = base + k;
        acc = tup_combine<IOP, SEGA, SEGB>(acc, item[k]);
        if (li < n) store(REVERSE ? n - 1 - li : li, acc, item[k]);
    }
}

template <int IOP, bool REVERSE, bool SEGA = false, bool SEGB = false, typename Load, typename Store>
int32_t scan_lookback(int64_t n, Load load, Store store, void *state_buf, cudaStream_t st) {
    if (n <= 0) return B200SURV_OK;
    const size_t ntiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    const ScanState s = scan_state_at(state_buf, n);
    const size_t words = 2 * ntiles * 3;
    unsigned ig = (unsigned)((words + 255) / 256);
    if (ig > 1184) ig = 1184;
    k_scan_state_init<<<ig, 256, 0, st>>>(s, words);
    k_scan_lookback<IOP, REVERSE, SEGA, SEGB, Load, Store><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(n, load, store, s);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

// ------------------------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // 4096 keys per tile; warp w owns keys [512 w, 512 w + 512)
constexpr int RS_RADIX = 256;

// per-tile digit histogram, stored digit-major: hist[digit][tile]
static __global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint32_t *__restrict__ keys, int64_t n, int shift, int ntiles, int *__restrict__ hist) {
    __shared__ int s_cnt[RS_RADIX];
    const int tile = blockIdx.x;
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)tile * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_cnt[(keys[i] >> shift) & (RS_RADIX - 1)], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + tile] = s_cnt[threadIdx.x];
}

// stable scatter: offs[digit][tile] = first output position of the tile's keys with that digit
static __global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, int shift, int ntiles,
             const long long *__restrict__ offs, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ int s_cnt[RS_THREADS / 32][RS_RADIX];       // running count of each digit within each warp's chunk
    __shared__ long long s_base[RS_THREADS / 32][RS_RADIX];
    const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < (RS_THREADS / 32) * RS_RADIX; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)tile * RS_TILE + (int64_t)warp * (RS_TILE / (RS_THREADS / 32));
    uint32_t k[RS_ROUNDS], v[RS_ROUNDS];
    int rank[RS_ROUNDS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool in = i < n;
        k[r] = in ? keys[i] : 0xffffffffu;
        v[r] = in ? vals[i] : 0u;
        const int d = in ? (int)((k[r] >> shift) & (RS_RADIX - 1)) : -1;
        // lanes of this round holding the same digit; out-of-range lanes form their own group and are ignored
        const unsigned peers = __match_any_sync(FULL, d);
        const int before = in ? s_cnt[warp][d] : 0;
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (in && (peers & lt) == 0) s_cnt[warp][d] = before + __popc(peers);  // the group's first lane updates the count
        __syncwarp();
    }
    __syncthreads();
    {   // digit t: exclusive prefix over the warps on top of the tile's global offset
        long long run = offs[(size_t)t * ntiles + tile];
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { s_base[w][t] = run; run += s_cnt[w][t]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) {
            const int d = (int)((k[r] >> shift) & (RS_RADIX - 1));
            const long long dst = s_base[warp][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = v[r];
        }
    }
}

struct RsLayout {
    size_t off_hist, off_offs, off_scan, total;
    int ntiles;
};
inline RsLayout rs_layout(int64_t n) {
    RsLayout L;
    L.ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    if (L.ntiles < 1) L.ntiles = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    L.off_hist = take((size_t)RS_RADIX * L.ntiles * sizeof(int));
    L.off_offs = take((size_t)RS_RADIX * L.ntiles * sizeof(long long));
    L.off_scan = take(scan_state_bytes((int64_t)RS_RADIX * L.ntiles));
    L.total = o;
    return L;
}
inline size_t radix_sort_temp_bytes(int64_t n) { return rs_layout(n).total; }

struct RsLoadHist {
    const int *hist;
    __device__ Tup operator()(int64_t p) const { Tup t; t.a = 0.0; t.b = 0.0; t.i = hist[p]; return t; }
};
struct RsStoreOffs {
    long long *offs;
    __device__ void operator()(int64_t p, const Tup &inc, const Tup &el) const { offs[p] = inc.i - el.i; }  // exclusive
};

// keys_in/vals_in are overwritten (ping-pong); the sorted pairs end up in keys_out/vals_out.  Bits [0, end_bit) of the
// keys are sorted (end_bit a multiple of 8, an even number of passes).
inline int32_t radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n,
                                int end_bit, void *temp, cudaStream_t st) {
    const RsLayout L = rs_layout(n);
    unsigned char *t8 = static_cast<unsigned char *>(temp);
    int *hist = reinterpret_cast<int *>(t8 + L.off_hist);
    long long *offs = reinterpret_cast<long long *>(t8 + L.off_offs);
    uint32_t *ka = keys_in, *va = vals_in, *kb = keys_out, *vb = vals_out;
    for (int shift = 0; shift < end_bit; shift += 8) {
        k_rs_hist<<<L.ntiles, RS_THREADS, 0, st>>>(ka, n, shift, L.ntiles, hist);
        const int32_t rc = scan_lookback<I_ADD, false>((int64_t)RS_RADIX * L.ntiles, RsLoadHist{hist}, RsStoreOffs{offs},
                                                       t8 + L.off_scan, st);
        if (rc) return rc;
        k_rs_scatter<<<L.ntiles, RS_THREADS, 0, st>>>(ka, va, n, shift, L.ntiles, offs, kb, vb);
        uint32_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;  // after an even number of passes the result is back in keys_in/vals_in: see the callers
}

// ------------------------------------------------------------------------------------------------ round 2
// (1) seg_scan: the look-back scan over a richer tuple -- three fp64 sums, each restarting at "group heads", at "segment
//     heads" or never, and a position combined with min / max.  Head flags travel in the integer, so the operator stays
//     associative:  (f1, v1) + (f2, v2) = (f1 | f2, f2 ? v2 : v1 + v2).  The SORTED Cox path runs entirely on it: tie
//     groups inside cohorts ("segments") packed back to back.
// (2) radix_sort_pairs2: the LSD radix sort with a scatter that is staged through shared memory (a tile's keys are first
//     put in digit order on chip, then written as runs: ~2 sectors per 16 keys instead of one sector per key) and a
//     key source functor, so that extra passes can sort by a key looked up through the value (the cohort id of a row).
struct Tup4 {
    double a, b, c;
    long long i;   // bits [0, 40): position + 1 (0 = none); bit 60: holds a group head; bit 61: holds a segment head
};
constexpr long long T4_POS = (1ll << 40) - 1, T4_GROUP = 1ll << 60, T4_SEG = 1ll << 61;
enum PosOp { P_NONE = 0, P_MIN = 1, P_MAX = 2 };

__device__ __forceinline__ Tup4 t4_identity() { Tup4 t; t.a = 0.0; t.b = 0.0; t.c = 0.0; t.i = 0; return t; }
// `x` precedes `y` in scan order.  RA / RB / RC: 0 = plain sum, 1 = restarts at group heads, 2 = restarts at segment heads.
template <int POP, int RA, int RB, int RC>
__device__ __forceinline__ Tup4 t4_combine(const Tup4 &x, const Tup4 &y) {
    Tup4 t;
    const bool hg = (y.i & T4_GROUP) != 0, hs = (y.i & T4_SEG) != 0;
    t.a = ((RA == 1 && hg) || (RA == 2 && hs)) ? y.a : x.a + y.a;
    t.b = ((RB == 1 && hg) || (RB == 2 && hs)) ? y.b : x.b + y.b;
    t.c = ((RC == 1 && hg) || (RC == 2 && hs)) ? y.c : x.c + y.c;
    const long long px = x.i & T4_POS, py = y.i & T4_POS;
    long long pos = 0;
    if (POP == P_MIN) pos = px == 0 ? py : (py == 0 ? px : (px < py ? px : py));
    if (POP == P_MAX) pos = px > py ? px : py;
    t.i = pos | ((x.i | y.i) & (T4_GROUP | T4_SEG));
    return t;
}
__device__ __forceinline__ Tup4 t4_shfl_up(const Tup4 &v, int d) {
    Tup4 t;
    t.a = __shfl_up_sync(FULL, v.a, d); t.b = __shfl_up_sync(FULL, v.b, d); t.c = __shfl_up_sync(FULL, v.c, d);
    t.i = __shfl_up_sync(FULL, v.i, d);
    return t;
}
__device__ __forceinline__ Tup4 t4_shfl(const Tup4 &v, int src) {
    Tup4 t;
    t.a = __shfl_sync(FULL, v.a, src); t.b = __shfl_sync(FULL, v.b, src); t.c = __shfl_sync(FULL, v.c, src);
    t.i = __shfl_sync(FULL, v.i, src);
    return t;
}
inline size_t seg_scan_state_bytes(int64_t n) {
    const size_t ntiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    return align_up(2 * ntiles * 4 * sizeof(unsigned long long), 256) + 256;
}
__device__ __forceinline__ void t4_publish(unsigned long long *w, const Tup4 &t) {
    scan_st(w, (unsigned long long)__double_as_longlong(t.a));
    scan_st(w + 1, (unsigned long long)__double_as_longlong(t.b));
    scan_st(w + 2, (unsigned long long)__double_as_longlong(t.c));
    scan_st(w + 3, (unsigned long long)t.i);
}
__device__ __forceinline__ bool t4_try_read(const unsigned long long *w, Tup4 &t) {
    const unsigned long long x = scan_ld(w), y = scan_ld(w + 1), z = scan_ld(w + 2), u = scan_ld(w + 3);
    t.a = __longlong_as_double((long long)x); t.b = __longlong_as_double((long long)y);
    t.c = __longlong_as_double((long long)z); t.i = (long long)u;
    return x != SCAN_EMPTY && y != SCAN_EMPTY && z != SCAN_EMPTY && u != SCAN_EMPTY;
}

// Inclusive scan of load(p); scan order = ascending p, or descending p when REVERSE.  store(p, inclusive, element); after its
// last element every thread calls store.finish() (block-wide reductions of per-thread state are allowed there: all threads
// of the block arrive).  grid = number of tiles, SCAN_THREADS threads; state: words[2][ntiles][4] preset to SCAN_EMPTY and a
// tile counter preset to 0 (k_scan_state_init over 8 * ntiles words).
template <int POP, bool REVERSE, int RA, int RB, int RC, typename Load, typename Store>
__global__ void __launch_bounds__(SCAN_THREADS, 3)   // <= 85 registers: three tiles per SM (at 148 it was one, 12 % occupancy)
k_seg_scan(int64_t n, Load load, Store store, unsigned long long *agg, unsigned long long *inc_w, unsigned *counter) {
    __shared__ Tup4 s_warp[SCAN_THREADS / 32];
    __shared__ Tup4 s_prefix;
    __shared__ unsigned s_tile;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SCAN_TILE + (int64_t)t * SCAN_ITEMS;
    Tup4 item[SCAN_ITEMS];
    Tup4 run = t4_identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t li = base + k;
        item[k] = li < n ? load(REVERSE ? n - 1 - li : li) : t4_identity();
        run = t4_combine<POP, RA, RB, RC>(run, item[k]);
    }
    Tup4 inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Tup4 u = t4_shfl_up(inc, d);
        if (lane >= d) inc = t4_combine<POP, RA, RB, RC>(u, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Tup4 wpre = t4_identity(), tile_agg = t4_identity();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) wpre = t4_combine<POP, RA, RB, RC>(wpre, s_warp[w]);
        tile_agg = t4_combine<POP, RA, RB, RC>(tile_agg, s_warp[w]);
    }
    Tup4 thread_excl = t4_shfl_up(inc, 1);
    if (lane == 0) thread_excl = t4_identity();
    thread_excl = t4_combine<POP, RA, RB, RC>(wpre, thread_excl);
    if (warp == 0) {  // decoupled look-back
        if (lane == 0) {
            if (tile == 0) t4_publish(inc_w, tile_agg);
            else t4_publish(agg + 4 * tile, tile_agg);
        }
        Tup4 prefix = t4_identity();
        int64_t look = tile - 1;
        for (int guard = 0; look >= 0 && guard < SCAN_SPIN_MAX; ++guard) {
            // lane l examines tile look - l: 2 = inclusive prefix available, 1 = aggregate only, 0 = nothing yet; the eight
            // words of both records are requested together (one round trip per poll)
            const int64_t q = look - lane;
            Tup4 v = t4_identity();
            int state = 2;  // lanes beyond tile 0 behave like "inclusive = identity"
            if (q >= 0) {
                Tup4 vi, va;
                const bool hi = t4_try_read(inc_w + 4 * q, vi), ha = t4_try_read(agg + 4 * q, va);
                state = hi ? 2 : (ha ? 1 : 0);
                v = hi ? vi : va;
            }
            const unsigned m_inc = __ballot_sync(FULL, state == 2), m_none = __ballot_sync(FULL, state == 0);
            const int first_inc = m_inc ? __ffs(m_inc) - 1 : 32, first_none = m_none ? __ffs(m_none) - 1 : 32;
            const bool done = first_inc < first_none;
            const int take = done ? first_inc + 1 : first_none;  // lanes [0, take) are combined (<= 32)
            // ordered tree reduction, farthest tile first: after step d lane i holds tiles (i + 2d - 1 .. i); five steps
            // instead of a 32-step serial chain of shuffles (that chain WAS the cost of a look-back hop)
            if (lane >= take) v = t4_identity();
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Tup4 u = t4_shfl(v, (lane + d) & 31);
                if (lane + d < 32) v = t4_combine<POP, RA, RB, RC>(u, v);
            }
            prefix = t4_combine<POP, RA, RB, RC>(t4_shfl(v, 0), prefix);
            if (done) break;
            look -= take;
        }
        if (lane == 0) {
            if (tile != 0) t4_publish(inc_w + 4 * tile, t4_combine<POP, RA, RB, RC>(prefix, tile_agg));
            s_prefix = prefix;
        }
    }
    __syncthreads();
    Tup4 acc = t4_combine<POP, RA, RB, RC>(s_prefix, thread_excl);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t li = base + k;
        acc = t4_combine<POP, RA, RB, RC>(acc, item[k]);
        if (li < n) store(REVERSE ? n - 1 - li : li, acc, item[k]);
    }
    store.finish();
}

template <int POP, bool REVERSE, int RA, int RB, int RC, typename Load, typename Store>
int32_t seg_scan(int64_t n, Load load, Store store, void *state_buf, cudaStream_t st) {
    if (n <= 0) return B200SURV_OK;
    const size_t ntiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    unsigned long long *agg = static_cast<unsigned long long *>(state_buf), *inc_w = agg + ntiles * 4;
    unsigned *counter = reinterpret_cast<unsigned *>(static_cast<unsigned char *>(state_buf) +
                                                     align_up(2 * ntiles * 4 * sizeof(unsigned long long), 256));
    ScanState s;
    s.agg = agg; s.inc = inc_w; s.counter = counter;
    const size_t words = 2 * ntiles * 4;
    unsigned ig = (unsigned)((words + 255) / 256);
    if (ig > 1184) ig = 1184;
    k_scan_state_init<<<ig, 256, 0, st>>>(s, words);
    k_seg_scan<POP, REVERSE, RA, RB, RC, Load, Store><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(n, load, store, agg, inc_w, counter);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

// ---- radix sort, round 2
struct KeyDirect {   // the sort key of element i is keys[i]
    __device__ uint32_t operator()(const uint32_t *keys, const uint32_t *, int64_t i) const { return keys[i]; }
};
struct KeyViaValue { // the sort key of element i is table[vals[i]] (e.g. the cohort of a row)
    const uint32_t *table;
    __device__ uint32_t operator()(const uint32_t *, const uint32_t *vals, int64_t i) const { return table[vals[i]]; }
};

template <typename KeyOf>
static __global__ void __launch_bounds__(RS_THREADS)
k_rs_hist2(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, KeyOf keyof, int64_t n, int shift, int ntiles,
           int *__restrict__ hist) {
    __shared__ int s_cnt[RS_RADIX];
    const int tile = blockIdx.x;
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)tile * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_cnt[(keyof(keys, vals, i) >> shift) & (RS_RADIX - 1)], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + tile] = s_cnt[threadIdx.x];
}

// stable scatter staged through shared memory: rank every key inside the tile (warp-level match ranking, warps in order),
// place the pairs in digit order on chip, then write each digit's run to its global position -- consecutive threads write
// consecutive addresses inside a run.
template <typename KeyOf>
static __global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter2(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, KeyOf keyof, int64_t n, int shift, int ntiles,
              const long long *__restrict__ offs, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ int s_cnt[RS_THREADS / 32][RS_RADIX];   // per warp: count of each digit, then its start inside the digit's run
    __shared__ int s_start[RS_RADIX];                  // first staged position of each digit
    __shared__ long long s_goff[RS_RADIX];             // global position of staged position 0 of each digit's run
    __shared__ uint32_t s_k[RS_TILE], s_v[RS_TILE];
    __shared__ unsigned char s_d[RS_TILE];
    const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < (RS_THREADS / 32) * RS_RADIX; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tbase = (int64_t)tile * RS_TILE;
    const int64_t base = tbase + (int64_t)warp * (RS_TILE / (RS_THREADS / 32));
    uint32_t k[RS_ROUNDS], v[RS_ROUNDS];
    int rank[RS_ROUNDS], dg[RS_ROUNDS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool in = i < n;
        k[r] = in ? keys[i] : 0xffffffffu;
        v[r] = in ? vals[i] : 0u;
        dg[r] = in ? (int)((keyof(keys, vals, i) >> shift) & (RS_RADIX - 1)) : -1;
        const unsigned peers = __match_any_sync(FULL, dg[r]);
        const int before = in ? s_cnt[warp][dg[r]] : 0;
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (in && (peers & lt) == 0) s_cnt[warp][dg[r]] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit t: total of the tile, per-warp starts inside the run
        int run = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { const int c = s_cnt[w][t]; s_cnt[w][t] = run; run += c; }
        // exclusive scan of the 256 digit totals (warp scans + 8 warp totals through s_start)
        int incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += u; }
        __shared__ int s_wt[RS_THREADS / 32];
        if (lane == 31) s_wt[warp] = incl;
        __syncthreads();
        int wpre = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) wpre += (w < warp) ? s_wt[w] : 0;
        const int start = wpre + incl - run;
        s_start[t] = start;
        s_goff[t] = offs[(size_t)t * ntiles + tile] - start;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        if (dg[r] >= 0) {
            const int pos = s_start[dg[r]] + s_cnt[warp][dg[r]] + rank[r];
            s_k[pos] = k[r]; s_v[pos] = v[r]; s_d[pos] = (unsigned char)dg[r];
        }
    }
    __syncthreads();
    const int64_t cnt = n - tbase < RS_TILE ? n - tbase : RS_TILE;
    for (int p = t; p < cnt; p += RS_THREADS) {
        const long long dst = s_goff[s_d[p]] + p;
        keys_out[dst] = s_k[p];
        vals_out[dst] = s_v[p];
    }
}

// One 8-bit pass: histogram, scan of the (digit, tile) matrix, staged scatter.
template <typename KeyOf>
inline int32_t radix_pass2(const uint32_t *ka, const uint32_t *va, uint32_t *kb, uint32_t *vb, KeyOf keyof, int64_t n, int shift,
                           void *temp, cudaStream_t st) {
    const RsLayout L = rs_layout(n);
    unsigned char *t8 = static_cast<unsigned char *>(temp);
    int *hist = reinterpret_cast<int *>(t8 + L.off_hist);
    long long *offs = reinterpret_cast<long long *>(t8 + L.off_offs);
    k_rs_hist2<KeyOf><<<L.ntiles, RS_THREADS, 0, st>>>(ka, va, keyof, n, shift, L.ntiles, hist);
    const int32_t rc = scan_lookback<I_ADD, false>((int64_t)RS_RADIX * L.ntiles, RsLoadHist{hist}, RsStoreOffs{offs},
                                                   t8 + L.off_scan, st);
    if (rc) return rc;
    k_rs_scatter2<KeyOf><<<L.ntiles, RS_THREADS, 0, st>>>(ka, va, keyof, n, shift, L.ntiles, offs, kb, vb);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

// Stable sort of (key, value) pairs by keys[0, end_bit) and then -- if seg_table != nullptr -- by seg_table[value]
// (seg_bits of it, a multiple of 8): pairs of one cohort end up contiguous, sorted by key inside the cohort.
// keys_in / vals_in are overwritten (ping-pong); the result is in keys_in / vals_in when the TOTAL number of passes is
// even, else in keys_out / vals_out: the function returns which through *in_first (1 = keys_in / vals_in).
inline int32_t radix_sort_pairs2(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n, int end_bit,
                                 const uint32_t *seg_table, int seg_bits, void *temp, cudaStream_t st, int *in_first) {
    uint32_t *ka = keys_in, *va = vals_in, *kb = keys_out, *vb = vals_out;
    int passes = 0;
    for (int shift = 0; shift < end_bit; shift += 8, ++passes) {
        const int32_t rc = radix_pass2(ka, va, kb, vb, KeyDirect{}, n, shift, temp, st);
        if (rc) return rc;
        uint32_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    if (seg_table != nullptr) {
        for (int shift = 0; shift < seg_bits; shift += 8, ++passes) {
            const int32_t rc = radix_pass2(ka, va, kb, vb, KeyViaValue{seg_table}, n, shift, temp, st);
            if (rc) return rc;
            uint32_t *tk = ka; ka = kb; kb = tk;
            uint32_t *tv = va; va = vb; vb = tv;
        }
    }
    *in_first = (passes % 2 == 0) ? 1 : 0;
    return B200SURV_OK;
}

}  // namespace sortscan
}  // namespace b200surv
