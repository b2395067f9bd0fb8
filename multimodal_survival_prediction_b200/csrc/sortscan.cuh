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
inline size_t radix_sort_temp_bytes_twokernel(int64_t n) { return rs_layout(n).total; }

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

// ------------------------------------------------------------------------------------------------ round 2b: one sweep
// LSD radix sort, ONE kernel per 8-bit pass ("onesweep"): a tile of 4096 pairs is ranked on chip (warp-level match ranking,
// warps in order), publishes its 256 digit counts and finds the counts of all earlier tiles by DECOUPLED LOOK-BACK -- one
// thread per digit, 64-bit words that carry their own state in the top two bits (0 = nothing yet, 1 = this tile's count,
// 2 = inclusive count up to this tile), so there is no flag to order against -- then writes its pairs as digit runs from
// shared memory.  The pairs are read once and written once per pass; the (digit, tile) count matrix, its scan and the
// second read of the keys of the two-kernel pass above are gone.  The global digit totals a pass needs up front are counted
// by the PREVIOUS pass (or by k_os_hist for the first).  Tile ids come from a ticket counter, so every tile a look-back
// waits for is already running.
constexpr int OS_THREADS = 256, OS_ITEMS = 16, OS_TILE = OS_THREADS * OS_ITEMS, OS_WARPS = OS_THREADS / 32;
constexpr unsigned long long OS_VAL = (1ull << 62) - 1;
constexpr int OS_MAX_PASSES = 8;

struct DigitFn {   // digit of a pair: 8 bits of the key, or of table[value] (e.g. the cohort of a row)
    const uint32_t *table;
    int shift;
};
__device__ __forceinline__ int os_digit(const DigitFn &f, uint32_t key, uint32_t val) {
    return (int)(((f.table ? f.table[val] : key) >> f.shift) & 255u);
}

static __global__ void __launch_bounds__(256)
k_os_hist(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, DigitFn f, unsigned *__restrict__ hist) {
    __shared__ unsigned s_cnt[256];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&s_cnt[os_digit(f, keys[i], f.table ? (vals ? vals[i] : (uint32_t)i) : 0u)], 1u);
    __syncthreads();
    if (s_cnt[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_cnt[threadIdx.x]);
}

// VIA_VAL: a digit (of this pass or of the next one) is looked up through the value, so the values are needed before the
// ranking; otherwise they are loaded late (fewer live registers while ranking).  vals == nullptr: the value of pair i is i.
template <bool VIA_VAL>
static __global__ void __launch_bounds__(OS_THREADS, 3)
k_onesweep(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, DigitFn cur, DigitFn nxt, int count_next,
           const unsigned *__restrict__ hist_cur, unsigned *__restrict__ hist_nxt, unsigned long long *state,
           unsigned long long *state_next, unsigned *ticket, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ unsigned s_cnt[OS_WARPS][256];   // per warp: count of each digit, then its start inside the digit's run
    __shared__ uint32_t s_k[OS_TILE], s_v[OS_TILE];
    __shared__ int s_start[256];                // first staged position of each digit
    __shared__ long long s_goff[256];           // global position of staged position 0 minus nothing: dst = s_goff[d] + p
    __shared__ unsigned s_next[256];
    __shared__ unsigned s_wt[2][OS_WARPS];
    __shared__ unsigned s_tile;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = t; i < OS_WARPS * 256; i += OS_THREADS) (&s_cnt[0][0])[i] = 0;
    s_next[t] = 0;
    __syncthreads();
    const int64_t tile = s_tile, tbase = tile * OS_TILE, base = tbase + (int64_t)warp * (OS_ITEMS * 32);
    uint32_t k[OS_ITEMS], v[OS_ITEMS];
    int rank[OS_ITEMS];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        k[r] = i < n ? keys[i] : 0xffffffffu;
    }
    if (VIA_VAL) {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const int64_t i = base + r * 32 + lane;
            v[r] = i < n ? (vals ? vals[i] : (uint32_t)i) : 0u;
        }
    }
    // ranking in three phases of 16 independent instructions each (match, add, shuffle), so that their latencies overlap:
    // lanes of a round with the same digit form a group; the group's first lane adds the group to the warp's running count
    // and hands the old value (= pairs of earlier rounds; shared atomics of one warp execute in program order) to the others
    const unsigned lt = (1u << lane) - 1u;
    unsigned peers[OS_ITEMS];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool in = i < n;
        const int d = in ? os_digit(cur, k[r], VIA_VAL ? v[r] : 0u) : 0;
        // lanes with the same digit, one ballot per digit bit (MATCH.ANY iterates over the distinct values of the warp:
        // ~30 of them for uniform digits, and it was the longest stall of the pass)
        unsigned m = __ballot_sync(FULL, in);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const unsigned bal = __ballot_sync(FULL, (d >> b) & 1);
            m &= ((d >> b) & 1) ? bal : ~bal;
        }
        peers[r] = in ? m : (1u << lane);
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        rank[r] = 0;
        if (i < n && (peers[r] & lt) == 0)
            rank[r] = (int)atomicAdd(&s_cnt[warp][os_digit(cur, k[r], VIA_VAL ? v[r] : 0u)], (unsigned)__popc(peers[r]));
    }
    if (count_next) {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const int64_t i = base + r * 32 + lane;
            if (i < n) atomicAdd(&s_next[os_digit(nxt, k[r], VIA_VAL ? v[r] : 0u)], 1u);
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r)
        rank[r] = __shfl_sync(FULL, rank[r], __ffs(peers[r]) - 1) + __popc(peers[r] & lt);
    __syncthreads();
    // digit t: per-warp starts inside the run, the tile's count (published at once), starts of the runs on chip
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) { const unsigned c = s_cnt[w][t]; s_cnt[w][t] = run; run += c; }
    unsigned long long *slot = state + (size_t)tile * 256 + t;
    state_next[(size_t)tile * 256 + t] = 0ull;   // the next pass's look-back words (its kernel starts after this one ends)
    if (tile > 0) scan_st(slot, (1ull << 62) | run);
    // the first window of the look-back is requested now and examined after the staging (its latency hides behind it)
    constexpr int LBW = 4;
    unsigned long long win[LBW];
    int64_t q = tile - 1;
#pragma unroll
    for (int j = 0; j < LBW; ++j) win[j] = q - j >= 0 ? scan_ld(state + (size_t)(q - j) * 256 + t) : (2ull << 62);
    long long gbase;
    int start;
    {
        unsigned inc_c = run, inc_h = hist_cur[t];
        const unsigned h = inc_h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned uc = __shfl_up_sync(FULL, inc_c, d), uh = __shfl_up_sync(FULL, inc_h, d);
            if (lane >= d) { inc_c += uc; inc_h += uh; }
        }
        if (lane == 31) { s_wt[0][warp] = inc_c; s_wt[1][warp] = inc_h; }
        __syncthreads();
        unsigned pre_c = 0, pre_h = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) { pre_c += (w < warp) ? s_wt[0][w] : 0u; pre_h += (w < warp) ? s_wt[1][w] : 0u; }
        start = (int)(pre_c + inc_c - run);
        gbase = (long long)(pre_h + inc_h - h);
        s_start[t] = start;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) {
            const int d = os_digit(cur, k[r], VIA_VAL ? v[r] : 0u);
            rank[r] += s_start[d] + (int)s_cnt[warp][d];
            s_k[rank[r]] = k[r];
        }
    }
    if (!VIA_VAL) {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const int64_t i = base + r * 32 + lane;
            v[r] = i < n ? (vals ? vals[i] : (uint32_t)i) : 0u;
        }
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) s_v[rank[r]] = v[r];
    }
    {   // look back over the earlier tiles' words for this digit, LBW words per round trip
        unsigned long long excl = 0;
        bool done = q < 0;
        while (!done) {
            int used = 0;
#pragma unroll
            for (int j = 0; j < LBW; ++j) {
                const unsigned fl = (unsigned)(win[j] >> 62);
                if (!done && used == j && fl != 0) {
                    excl += win[j] & OS_VAL; ++used;
                    if (fl == 2) done = true;
                }
            }
            q -= used;
            if (q < 0) done = true;
            if (!done) {
#pragma unroll
                for (int j = 0; j < LBW; ++j) win[j] = q - j >= 0 ? scan_ld(state + (size_t)(q - j) * 256 + t) : (2ull << 62);
            }
        }
        scan_st(slot, (2ull << 62) | (excl + run));
        s_goff[t] = gbase + (long long)excl - start;
    }
    __syncthreads();
    const int cnt = (int)(n - tbase < OS_TILE ? n - tbase : OS_TILE);
#pragma unroll 4
    for (int p = t; p < cnt; p += OS_THREADS) {
        const uint32_t key = s_k[p], val = s_v[p];
        const long long dst = s_goff[os_digit(cur, key, val)] + p;
        keys_out[dst] = key;
        vals_out[dst] = val;
    }
    if (count_next && s_next[t]) atomicAdd(&hist_nxt[t], s_next[t]);
}

// The same pass when both digits (this pass's and the next one's) are bits of the KEY and n < 2^31: the common case (every
// pass of a single cohort / of the C-index; all but the last passes of packed cohorts).  k_onesweep above is issue-bound at
// ~220 SASS instructions per pair (ncu r2_v3: 115 M warp instructions per pass at 16.7M pairs); here the per-pair work is
// 32-bit throughout: no bounds tests on full tiles (a partial tile pads with key 0xffffffff -- the pads rank last in digit
// 255 and are taken out of the published count), one 8-byte staged word per pair, 32-bit run offsets modulo 2^32.
// lanes whose bit b of d DIFFERS from this lane's: ballot and one predicated complement.  PTX, so that the eight bit tests
// of a digit become one R2P (the C form `m &= bit ? bal : ~bal` compiles to five or six instructions per bit).
__device__ __forceinline__ unsigned os_mismatch_bit(unsigned d, int b) {
    unsigned x;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.b32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "@p not.b32 %0, %0;\n\t}"
        : "=r"(x) : "r"(d), "r"(1u << b));
    return x;
}
// lanes of the warp whose low 8 bits of d equal this lane's (three-input ORs: 8 + 8 + 4 instructions and the R2P)
__device__ __forceinline__ unsigned os_match8(unsigned d) {
    const unsigned x0 = os_mismatch_bit(d, 0), x1 = os_mismatch_bit(d, 1), x2 = os_mismatch_bit(d, 2), x3 = os_mismatch_bit(d, 3),
                   x4 = os_mismatch_bit(d, 4), x5 = os_mismatch_bit(d, 5), x6 = os_mismatch_bit(d, 6), x7 = os_mismatch_bit(d, 7);
    const unsigned a = x0 | x1 | x2, b = x3 | x4 | x5, c = x6 | x7 | a;
    return ~(b | c);
}
template <int MINB>
static __global__ void __launch_bounds__(OS_THREADS, MINB)
k_onesweep_key(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, unsigned n, int shift, int shift_next,
               int count_next, const unsigned *__restrict__ hist_cur, unsigned *__restrict__ hist_nxt, unsigned long long *state,
               unsigned long long *state_next, unsigned *ticket, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ unsigned s_cnt[OS_WARPS][256];   // per warp: count of each digit, then staged start of the warp's run
    __shared__ uint32_t s_k[OS_TILE], s_v[OS_TILE];
    __shared__ unsigned s_goff[256];            // global position minus staged position of a digit's run (mod 2^32)
    __shared__ unsigned s_next[256];
    __shared__ unsigned s_wt[2][OS_WARPS];
    __shared__ unsigned s_tile;
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) s_cnt[w][t] = 0;
    s_next[t] = 0;
    __syncthreads();
    const unsigned tile = s_tile, tbase = tile * OS_TILE, wbase = tbase + warp * (OS_ITEMS * 32) + lane;
    const unsigned cnt = n - tbase < (unsigned)OS_TILE ? n - tbase : (unsigned)OS_TILE;
    const bool full = cnt == (unsigned)OS_TILE;
    uint32_t k[OS_ITEMS];
    if (full) {
        const uint32_t *kp = keys + wbase;
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) k[r] = kp[r * 32];
    } else {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) k[r] = wbase + r * 32 < n ? keys[wbase + r * 32] : 0xffffffffu;
    }
    // ranking: lanes of a round with the same digit form a group (one ballot per digit bit); the group's first lane adds the
    // group to the warp's running count (shared atomics of one warp execute in program order) and hands the old value on
    const unsigned lt = (1u << lane) - 1u;
    unsigned peers[OS_ITEMS], rank[OS_ITEMS];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        peers[r] = os_match8(k[r] >> shift);
    }
    unsigned *my_cnt = &s_cnt[warp][0];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        rank[r] = 0;
        if ((peers[r] & lt) == 0) rank[r] = atomicAdd(my_cnt + ((k[r] >> shift) & 255u), (unsigned)__popc(peers[r]));
    }
    if (count_next) {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) atomicAdd(&s_next[(k[r] >> shift_next) & 255u], 1u);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r)
        rank[r] = __shfl_sync(FULL, rank[r], __ffs(peers[r]) - 1) + __popc(peers[r] & lt);
    __syncthreads();
    // digit t: per-warp counts -> the tile's count (published at once), staged starts of the runs
    unsigned c[OS_WARPS], run = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) { c[w] = run; run += s_cnt[w][t]; }
    const unsigned pads = (unsigned)OS_TILE - cnt;          // all in digit 255, behind the real pairs
    const unsigned run_pub = t == 255u ? run - pads : run;
    unsigned long long *slot = state + (size_t)tile * 256 + t;
    state_next[(size_t)tile * 256 + t] = 0ull;   // the next pass's look-back words (its kernel starts after this one ends)
    if (tile > 0) scan_st(slot, (1ull << 62) | run_pub);
    constexpr int LBW = 4;
    unsigned long long win[LBW];
    int q = (int)tile - 1;
#pragma unroll
    for (int j = 0; j < LBW; ++j) win[j] = q - j >= 0 ? scan_ld(state + (size_t)(q - j) * 256 + t) : (2ull << 62);
    unsigned gbase, start;
    {
        unsigned inc_c = run, inc_h = hist_cur[t];
        const unsigned h = inc_h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned uc = __shfl_up_sync(FULL, inc_c, d), uh = __shfl_up_sync(FULL, inc_h, d);
            if (lane >= (unsigned)d) { inc_c += uc; inc_h += uh; }
        }
        if (lane == 31u) { s_wt[0][warp] = inc_c; s_wt[1][warp] = inc_h; }
        __syncthreads();
        unsigned pre_c = 0, pre_h = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) { pre_c += ((unsigned)w < warp) ? s_wt[0][w] : 0u; pre_h += ((unsigned)w < warp) ? s_wt[1][w] : 0u; }
        start = pre_c + inc_c - run;
        gbase = pre_h + inc_h - h;
    }
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) s_cnt[w][t] = start + c[w];
    __syncthreads();
    // keys into their staged places; the values are loaded only now (the keys' registers are free) and land after the
    // look-back, which hides their latency
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        rank[r] += my_cnt[(k[r] >> shift) & 255u];
        s_k[rank[r]] = k[r];
    }
    uint32_t v[OS_ITEMS];
    if (vals == nullptr) {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) v[r] = wbase + r * 32;
    } else if (full) {
        const uint32_t *vp = vals + wbase;
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) v[r] = vp[r * 32];
    } else {
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) v[r] = wbase + r * 32 < n ? vals[wbase + r * 32] : 0u;
    }
    {   // look back over the earlier tiles' words for this digit, LBW words per round trip
        unsigned excl = 0;
        bool done = q < 0;
        while (!done) {
            int used = 0;
#pragma unroll
            for (int j = 0; j < LBW; ++j) {
                const unsigned fl = (unsigned)(win[j] >> 62);
                if (!done && used == j && fl != 0) {
                    excl += (unsigned)win[j]; ++used;
                    if (fl == 2) done = true;
                }
            }
            q -= used;
            if (q < 0) done = true;
            if (!done) {
#pragma unroll
                for (int j = 0; j < LBW; ++j) win[j] = q - j >= 0 ? scan_ld(state + (size_t)(q - j) * 256 + t) : (2ull << 62);
            }
        }
        scan_st(slot, (2ull << 62) | (unsigned long long)(excl + run_pub));
        s_goff[t] = gbase + excl - start;
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) s_v[rank[r]] = v[r];
    __syncthreads();
    if (full) {
#pragma unroll
        for (int i = 0; i < OS_ITEMS; ++i) {
            const unsigned p = t + i * OS_THREADS;
            const uint32_t key = s_k[p];
            const unsigned dst = s_goff[(key >> shift) & 255u] + p;
            keys_out[dst] = key;
            vals_out[dst] = s_v[p];
        }
    } else {
        for (unsigned p = t; p < cnt; p += OS_THREADS) {
            const uint32_t key = s_k[p];
            const unsigned dst = s_goff[(key >> shift) & 255u] + p;
            keys_out[dst] = key;
            vals_out[dst] = s_v[p];
        }
    }
    if (count_next) {
        const unsigned cn = (t == ((0xffffffffu >> shift_next) & 255u)) ? s_next[t] - pads : s_next[t];
        if (cn) atomicAdd(&hist_nxt[t], cn);
    }
}

struct OsLayout {
    size_t off_hist, off_ticket, off_state, off_state2, total;
    int ntiles;
};
inline OsLayout os_layout(int64_t n) {
    OsLayout L;
    L.ntiles = (int)((n + OS_TILE - 1) / OS_TILE);
    if (L.ntiles < 1) L.ntiles = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    L.off_hist = take((size_t)OS_MAX_PASSES * 256 * sizeof(unsigned));
    L.off_ticket = take((size_t)OS_MAX_PASSES * sizeof(unsigned));
    L.off_state = take((size_t)L.ntiles * 256 * sizeof(unsigned long long));    // look-back words of the even passes ...
    L.off_state2 = take((size_t)L.ntiles * 256 * sizeof(unsigned long long));   // ... and of the odd ones (each pass zeroes the other set)
    L.total = o;
    return L;
}
inline size_t onesweep_temp_bytes(int64_t n) { return os_layout(n).total; }
inline size_t radix_sort_temp_bytes(int64_t n) { return os_layout(n).total; }

// Stable sort of (key, value) pairs by keys[0, end_bit) and then -- if seg_table != nullptr -- by seg_table[value]
// (seg_bits of it, a multiple of 8): pairs of one cohort end up contiguous, sorted by key inside the cohort.
// vals_in == nullptr: the value of pair i is i (nothing is read for the first pass).  keys_in / vals_in are overwritten
// (ping-pong; vals_in must still be a buffer of n values); the result is in keys_in / vals_in when the TOTAL number of
// passes is even, else in keys_out / vals_out: the function returns which through *in_first (1 = keys_in / vals_in).
// The producer of the keys may count the first pass's digits itself (low 8 bits of the key) and save the sort one read of the
// keys: radix_sort_prepare() zeroes the sort's counters and returns the 256 totals to add into; the sort is then called
// with first_hist_done = true.
inline int32_t radix_sort_prepare(int64_t n, void *temp, cudaStream_t st, unsigned **hist0);
inline int32_t radix_sort_pairs2(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n, int end_bit,
                                 const uint32_t *seg_table, int seg_bits, void *temp, cudaStream_t st, int *in_first,
                                 bool vals_are_iota = false, bool first_hist_done = false) {
    const OsLayout L = os_layout(n);
    unsigned char *t8 = static_cast<unsigned char *>(temp);
    unsigned *hist = reinterpret_cast<unsigned *>(t8 + L.off_hist), *ticket = reinterpret_cast<unsigned *>(t8 + L.off_ticket);
    unsigned long long *state_ab[2] = {reinterpret_cast<unsigned long long *>(t8 + L.off_state),
                                       reinterpret_cast<unsigned long long *>(t8 + L.off_state2)};
    DigitFn fn[OS_MAX_PASSES + 1];
    int passes = 0;
    for (int shift = 0; shift < end_bit; shift += 8) fn[passes++] = DigitFn{nullptr, shift};
    if (seg_table != nullptr)
        for (int shift = 0; shift < seg_bits; shift += 8) fn[passes++] = DigitFn{seg_table, shift};
    if (passes > OS_MAX_PASSES) { set_error("radix sort: %d passes > %d", passes, OS_MAX_PASSES); return B200SURV_BAD_ARG; }
    *in_first = (passes % 2 == 0) ? 1 : 0;
    if (n <= 0 || passes == 0) return B200SURV_OK;
    if (!first_hist_done) {
        B200_CHECK_CUDA(cudaMemsetAsync(t8 + L.off_hist, 0, L.off_state2 - L.off_hist, st));   // digit totals, tickets, first pass's look-back words
        int hg = (int)((n + 255) / 256);
        if (hg > 8 * num_sms()) hg = 8 * num_sms();
        const uint32_t *va0 = vals_are_iota ? nullptr : vals_in;
        k_os_hist<<<hg, 256, 0, st>>>(keys_in, va0, n, fn[0], hist);
    }
    uint32_t *ka = keys_in, *va = vals_in, *kb = keys_out, *vb = vals_out;
    for (int p = 0; p < passes; ++p) {
        unsigned long long *state = state_ab[p & 1], *state_next = state_ab[(p + 1) & 1];
        const bool has_next = p + 1 < passes;
        const DigitFn nxt = has_next ? fn[p + 1] : DigitFn{nullptr, 0};
        const uint32_t *vsrc = (p == 0 && vals_are_iota) ? nullptr : va;
        const bool via_val = fn[p].table != nullptr || (has_next && nxt.table != nullptr);
        if (via_val)
            k_onesweep<true><<<L.ntiles, OS_THREADS, 0, st>>>(ka, vsrc, n, fn[p], nxt, has_next ? 1 : 0, hist + p * 256,
                                                              hist + (p + 1) * 256, state, state_next, ticket + p, kb, vb);
        else if (n < (1ll << 31))   // 2, 3 or 4 CTAs per SM time alike (121-126 us per pass at 16.7M pairs): the LSU pipe is the bound
            k_onesweep_key<3><<<L.ntiles, OS_THREADS, 0, st>>>(ka, vsrc, (unsigned)n, fn[p].shift, nxt.shift, has_next ? 1 : 0,
                                                               hist + p * 256, hist + (p + 1) * 256, state, state_next, ticket + p, kb, vb);
        else
            k_onesweep<false><<<L.ntiles, OS_THREADS, 0, st>>>(ka, vsrc, n, fn[p], nxt, has_next ? 1 : 0, hist + p * 256,
                                                               hist + (p + 1) * 256, state, state_next, ticket + p, kb, vb);
        uint32_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

inline int32_t radix_sort_prepare(int64_t n, void *temp, cudaStream_t st, unsigned **hist0) {
    const OsLayout L = os_layout(n);
    unsigned char *t8 = static_cast<unsigned char *>(temp);
    B200_CHECK_CUDA(cudaMemsetAsync(t8 + L.off_hist, 0, L.off_state2 - L.off_hist, st));
    *hist0 = reinterpret_cast<unsigned *>(t8 + L.off_hist);
    return B200SURV_OK;
}

}  // namespace sortscan
}  // namespace b200surv
