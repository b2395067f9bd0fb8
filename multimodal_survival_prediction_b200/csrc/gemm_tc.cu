// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// This is the dense contraction of the fusion head: the Linear layers of PartialModalityNet /
// MultiModalSurvivalNet (scripts/training/partial_modality_training.py:196-232) and their
// weight/input gradients.  One kernel, two operand layouts per side:
//   K-major  operand: row-major [rows][K]  (activations [B][K], weights [out][in])
//   MN-major operand: row-major [K][cols]  (the same buffers read "transposed": dW = dY^T X needs
//                     dY as [K=B][M=out] and X as [K=B][N=in] -- no transposed copies are made)
//   C[M][N] (fp32, + optional bf16 copy) = A * B (+ bias[n]) (ReLU)
//
// Structure (one 128 x BN output tile per CTA, BN = 128 / 192 / 256 columns of TMEM; 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier tx-count
//   warp 1      allocates 128 TMEM columns, issues tcgen05.mma.cta_group::1.kind::f16 (one elected
//               lane), tcgen05.commit frees smem stages / signals the accumulator
//   warps 2..9  epilogue: tcgen05.ld (32 lanes x 32 columns at a time) -> bias/ReLU -> shared-memory staging -> global
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200surv {
long long *g_gemm_trace = nullptr;
namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int GEMM_THREADS = 64 + 8 * 32;  // TMA warp, MMA warp, eight epilogue warps
constexpr int TILE_BYTES = BM * BK * 2;  // A operand: 16 KB per stage; B operand: BN * BK * 2
// Tile widths.  128: the default (most CTAs for the small GEMMs).  256: rna_encoder.0 forward -- a 128x128 tile loads
// 32 KB per 2.1 MFLOP, 328 MB of L2->SM traffic for the whole GEMM, which is what bounds it (36 us against 15 us of
// tensor work); 128x256 moves 246 MB.  192: the rna_encoder.0 weight gradient, 4 x 27 = 108 tiles in ONE wave where
// 128-wide tiles made 160 CTAs = two waves on 148 SMs.
template <int BN_> struct TileCfg {
    static constexpr int STAGES = BN_ == 256 ? 4 : 5;
    static constexpr int TILE_B = BN_ * BK * 2;
    static constexpr int TMEM_COLS = BN_ <= 128 ? 128 : 256;
    static constexpr size_t SMEM = (size_t)STAGES * (TILE_BYTES + TILE_B) + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr int MBAR_SPIN_MAX = 1 << 27;  // a broken pipeline must trap, not hang the GPU

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < MBAR_SPIN_MAX; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1, [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, majorness, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Epilogue {
    float *c;               // [M][ldc] fp32 (nullable)
    __nv_bfloat16 *c_bf16;  // [M][ldc_bf16] bf16 copy (nullable)
    const float *bias;      // [N] (nullable)
    int64_t ldc, ldc_bf16;
    int relu;
    int kb_per;             // split-K: k-blocks per grid.z slice (slice z writes c + z * split_stride)
    int64_t split_stride;
    int staged;             // exactly one output, 16-byte aligned: rows are staged in shared memory and stored coalesced
    long long *trace;       // diagnostics (b200surv_debug_gemm_trace): globaltimer stamps of CTA (0,0,0), or null
};
__device__ __forceinline__ long long gtimer() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}


// tcgen05.ld of 32 lanes x 32 columns WITHOUT the wait (the caller overlaps it with the previous chunk's work)
__device__ __forceinline__ void tmem_ld_32x32_async(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue of one 128 x BN accumulator tile.  EIGHT warps: quarter q = warp % 4 owns TMEM lanes [32 q, 32 q + 32) = 32 rows
// of the tile (a hardware rule), and the two warps of a quarter split its columns (half = 0 / 1).  With ONE output the
// warp stages its 32 rows x BN / 2 columns in the pipeline's shared memory (free: every MMA has retired) and stores whole
// rows -- after tcgen05.ld a lane holds a ROW, so a direct store touches 32 rows per instruction (32 L1 tag cycles each;
// that alone cost 25 us on the 512 x 5005 weight gradient).  ep.staged: 1 = 16-byte vectors (aligned outputs), 2 = 4-byte
// coalesced stores (fp32 rows that are not 16-byte aligned, e.g. ldc = 5005), 0 = direct stores (two outputs).
// The TMEM loads are software-pipelined: chunk i + 1 is in flight while chunk i is converted and staged.
constexpr int EPI_WARPS = 8;
template <int BN>
__device__ __forceinline__ void run_epilogue(const Epilogue &ep, uint32_t tmem_base, int tile_m, int tile_n, int M, int N,
                                             unsigned char *tiles_a, int q, int half, int lane) {
    constexpr int CW = BN / 2;                       // columns of this warp
    static_assert(CW % 32 == 0 || CW == 96, "column split");
    const int row0 = tile_m * BM + q * 32;
    const int cbeg = half * CW;
    if (row0 >= M || tile_n * BN + cbeg >= N) return;  // (warp-uniform) nothing of this warp's block is inside the matrix
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
    constexpr int NCH = CW / 32;
    uint32_t va[32], vb[32];
    tmem_ld_32x32_async(taddr, va);
    tmem_ld_wait();
    if (ep.staged) {
        const int ES = ep.c != nullptr ? 4 : 2, pitch = CW * ES + 16;
        unsigned char *stg = tiles_a + (size_t)(q * 2 + half) * (32 * (CW * 4 + 16));
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            uint32_t(&cur)[32] = (ch & 1) ? vb : va;
            uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
            if (ch + 1 < NCH) tmem_ld_32x32_async(taddr + (uint32_t)(32 * (ch + 1)), nxt);
            const int c = 32 * ch, col0 = tile_n * BN + cbeg + c;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(cur[j]);
                if (ep.bias != nullptr && col0 + j < N) x += __ldg(ep.bias + col0 + j);
                if (ep.relu) x = fmaxf(x, 0.f);
                f[j] = x;
            }
            if (ES == 4) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(stg + lane * pitch + (c + j) * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]),
                                   p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t *>(&p0); u.y = *reinterpret_cast<uint32_t *>(&p1);
                    u.z = *reinterpret_cast<uint32_t *>(&p2); u.w = *reinterpret_cast<uint32_t *>(&p3);
                    *reinterpret_cast<uint4 *>(stg + lane * pitch + (c + j) * 2) = u;
                }
            }
            if (ch + 1 < NCH) tmem_ld_wait();
        }
        __syncwarp();
        const int gc0 = tile_n * BN + cbeg;
        if (ep.staged == 1) {
            const int vpr = CW * ES / 16, per_vec = 16 / ES;  // 16-byte vectors per staged row; consecutive lanes take
#pragma unroll 4
            for (int idx = lane; idx < 32 * vpr; idx += 32) {  // consecutive vectors of a row: whole-row stores
                const int k = idx / vpr, vec = idx - k * vpr;
                const int grow = row0 + k, gcol = gc0 + vec * per_vec;
                if (grow < M && gcol < N) {                     // N is a multiple of per_vec: whole vector inside
                    const uint4 u = *reinterpret_cast<const uint4 *>(stg + k * pitch + vec * 16);
                    unsigned char *dst = ES == 4
                        ? reinterpret_cast<unsigned char *>(ep.c + (int64_t)blockIdx.z * ep.split_stride + (int64_t)grow * ep.ldc + gcol)
                        : reinterpret_cast<unsigned char *>(ep.c_bf16 + (int64_t)grow * ep.ldc_bf16 + gcol);
                    *reinterpret_cast<uint4 *>(dst) = u;
                }
            }
        } else {  // fp32 rows without 16-byte alignment: lanes along the columns, 128 contiguous bytes per store
            float *cbase = ep.c + (int64_t)blockIdx.z * ep.split_stride;
            const int kmax = min(32, M - row0);
            for (int k = 0; k < kmax; ++k) {
                float *dst = cbase + (int64_t)(row0 + k) * ep.ldc + gc0;
                const float *src = reinterpret_cast<const float *>(stg + k * pitch);
#pragma unroll
                for (int j = lane; j < CW; j += 32)
                    if (gc0 + j < N) dst[j] = src[j];
            }
        }
        return;
    }
    const int row = row0 + lane;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        uint32_t(&cur)[32] = (ch & 1) ? vb : va;
        uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
        if (ch + 1 < NCH) tmem_ld_32x32_async(taddr + (uint32_t)(32 * (ch + 1)), nxt);
        const int col0 = tile_n * BN + cbeg + 32 * ch;
        if (row < M && col0 < N) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(cur[j]);
                if (ep.bias != nullptr && col0 + j < N) x += __ldg(ep.bias + col0 + j);
                if (ep.relu) x = fmaxf(x, 0.f);
                f[j] = x;
            }
            if (ep.c != nullptr) {
                float *dst = ep.c + (int64_t)blockIdx.z * ep.split_stride + (int64_t)row * ep.ldc + col0;
                if (col0 + 32 <= N && (ep.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(ep.c) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                } else {
                    for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = f[j];
                }
            }
            if (ep.c_bf16 != nullptr) {
                __nv_bfloat16 *dst = ep.c_bf16 + (int64_t)row * ep.ldc_bf16 + col0;
                if (col0 + 32 <= N && (ep.ldc_bf16 & 7) == 0 && ((reinterpret_cast<uintptr_t>(ep.c_bf16) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]),
                                       p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                        uint4 u;
                        u.x = *reinterpret_cast<uint32_t *>(&p0); u.y = *reinterpret_cast<uint32_t *>(&p1);
                        u.z = *reinterpret_cast<uint32_t *>(&p2); u.w = *reinterpret_cast<uint32_t *>(&p3);
                        *reinterpret_cast<uint4 *>(dst + j) = u;
                    }
                } else {
                    for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = __float2bfloat16_rn(f[j]);
                }
            }
        }
        if (ch + 1 < NCH) tmem_ld_wait();
    }
}

// A_MN / B_MN: operand is stored [K][cols] (MN-major) instead of [rows][K] (K-major)
template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N, int K,
             Epilogue ep) {
    pdl_trigger();   // the next kernel of the chain may set itself up while this one runs
    constexpr int STAGES = TileCfg<BN>::STAGES, TILE_B = TileCfg<BN>::TILE_B, TMEM_COLS = TileCfg<BN>::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *tiles_a = smem;
    unsigned char *tiles_b = smem + (size_t)STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * (TILE_BYTES + TILE_B));
    // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready; then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.y, tile_n = blockIdx.x;
    const int kb0 = blockIdx.z * ep.kb_per, kb1 = min((K + BK - 1) / BK, kb0 + ep.kb_per);  // this slice's k-blocks

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + STAGES + s), 1); }
        mbar_init(smem_u32(bars + 2 * STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // barriers, TMEM and tensor maps are set up: now the previous kernel's results are needed

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + STAGES + stage), phase ^ 1);
                const uint32_t full = smem_u32(bars + stage);
                mbar_expect_tx(full, TILE_BYTES + TILE_B);
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_B);
                if (A_MN) {  // two [64 k][64 m] boxes
                    tma_load_2d(sa, &map_a, tile_m * BM, kb * BK, full);
                    tma_load_2d(sa + TILE_BYTES / 2, &map_a, tile_m * BM + 64, kb * BK, full);
                } else {     // one [128 m][64 k] box
                    tma_load_2d(sa, &map_a, kb * BK, tile_m * BM, full);
                }
                if (B_MN) {  // BN / 64 boxes of [64 k][64 n], 8 KB apart
#pragma unroll
                    for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * (TILE_BYTES / 2), &map_b, tile_n * BN + 64 * i, kb * BK, full);
                } else {     // one [BN n][64 k] box
                    tma_load_2d(sb, &map_b, kb * BK, tile_n * BN, full);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(A_MN, B_MN, BN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + stage), phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_B);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // K-major : 8-row groups 1024 B apart (SBO), 16 k-elements = 32 B along the row
                    // MN-major: 64-element atoms 8192 B apart (LBO), 8-k groups 1024 B apart (SBO),
                    //           16 k = 2 groups = 2048 B
                    const uint64_t da = A_MN ? make_desc(sa + k * 2048, TILE_BYTES / 2, 1024) : make_desc(sa + k * 32, 16, 1024);
                    const uint64_t db = B_MN ? make_desc(sb + k * 2048, TILE_BYTES / 2, 1024) : make_desc(sb + k * 32, 16, 1024);
                    umma_f16(tmem_base, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(bars + STAGES + stage));  // frees this smem stage when the MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(smem_u32(bars + 2 * STAGES));          // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
        mbar_wait(smem_u32(bars + 2 * STAGES), 0);
        tcgen05_fence_after();
        run_epilogue<BN>(ep, tmem_base, tile_m, tile_n, M, N, tiles_a, warp & 3, (warp - 2) >> 2, lane);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================ 2-CTA pairs (cta_group::2): 256 x 256 tiles
// What bounds the one-CTA kernel above on the two big GEMMs is not the tensor pipe and not L2: one SM takes in about
// 46 bytes per clock through TMA (ncu: 331 MB in 36 us over 128 SMs = 37 B/clk/SM with the tensor pipe 37 % busy, L2 at
// 20 %), and a 128 x 128 x 64 k-block is 32 KB for 271 clocks of tensor work.  A CTA PAIR on the two SMs of a TPC
// computes a 256 x 256 tile with ONE tcgen05.mma.cta_group::2 stream issued by the leader: each CTA still loads 32 KB
// per k-block (its 128 rows of A, its 128 columns of B) but does twice the math on it (128 FLOP per byte).
//   grid (2 * ceil(M / 256), ceil(N / 256), k-splits), cluster (2, 1, 1); CTA rank r of a pair owns rows
//   [256 tx + 128 r, +128) -- its TMEM holds those 128 rows x 256 columns -- and columns [256 ty + 128 r, +128) of B.
//   TMA loads of both CTAs complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2 with the
//   leader's barrier address); tcgen05.commit ... .multicast::cluster releases the stage in both CTAs and signals both
//   epilogues.
constexpr int STAGES2 = 6;
constexpr size_t GEMM2_SMEM = (size_t)STAGES2 * 2 * TILE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair(bool a_mn, bool b_mn) {  // M = 256, N = 256
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(256 >> 3) << 17) |
           ((uint32_t)(256 >> 4) << 24);
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N, int K,
              Epilogue ep) {
    pdl_trigger();   // the next kernel of the chain may set itself up while this one runs
    constexpr int STAGES = STAGES2, BN = 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *tiles_a = smem;
    unsigned char *tiles_b = smem + (size_t)STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * 2 * TILE_BYTES);  // full[S], empty[S], accumulator
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tile_m = blockIdx.x;            // 128-row tile of THIS CTA (the pair covers tiles 2 * (x / 2), + 1)
    const int tile_n = blockIdx.y;            // 256-column tile of the pair
    const int kb0 = blockIdx.z * ep.kb_per, kb1 = min((K + BK - 1) / BK, kb0 + ep.kb_per);
    const bool tr = ep.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (tr && threadIdx.x == 0) ep.trace[0] = gtimer();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + STAGES + s), 1); }
        mbar_init(smem_u32(bars + 2 * STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // both CTAs' barriers exist before anybody signals across the pair
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // barriers, TMEM and tensor maps are set up: now the previous kernel's results are needed

    if (warp == 0) {
        // ===== TMA producer (one per CTA): own A rows, own half of the B columns; completion on the leader's barrier =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int n0 = tile_n * BN + (int)rank * 128;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + STAGES + stage), phase ^ 1);
                const uint32_t full_local = smem_u32(bars + stage);
                if (leader) mbar_expect_tx(full_local, 4 * TILE_BYTES);  // A and B halves of both CTAs
                const uint32_t full = map_to_cta(full_local, 0);
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_BYTES);
                if (A_MN) {
                    tma_load_2d_pair(sa, &map_a, tile_m * BM, kb * BK, full);
                    tma_load_2d_pair(sa + TILE_BYTES / 2, &map_a, tile_m * BM + 64, kb * BK, full);
                } else {
                    tma_load_2d_pair(sa, &map_a, kb * BK, tile_m * BM, full);
                }
                if (B_MN) {
                    tma_load_2d_pair(sb, &map_b, n0, kb * BK, full);
                    tma_load_2d_pair(sb + TILE_BYTES / 2, &map_b, n0 + 64, kb * BK, full);
                } else {
                    tma_load_2d_pair(sb, &map_b, kb * BK, n0, full);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's elected lane, for the pair =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_pair(A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            if (tr) ep.trace[1] = gtimer();
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + stage), phase);
                tcgen05_fence_after();
                if (tr && kb - kb0 < 24) ep.trace[8 + (kb - kb0)] = gtimer();
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t da = A_MN ? make_desc(sa + k * 2048, TILE_BYTES / 2, 1024) : make_desc(sa + k * 32, 16, 1024);
                    const uint64_t db = B_MN ? make_desc(sb + k * 2048, TILE_BYTES / 2, 1024) : make_desc(sb + k * 32, 16, 1024);
                    umma_f16_pair(tmem_base, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit_pair(smem_u32(bars + STAGES + stage));  // frees the stage in BOTH CTAs
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(smem_u32(bars + 2 * STAGES));          // accumulators complete, both CTAs
            if (tr) ep.trace[2] = gtimer();
        }
    } else {
        // ===== epilogue: each CTA drains its own 128 rows x 256 columns =====
        mbar_wait(smem_u32(bars + 2 * STAGES), 0);
        tcgen05_fence_after();
        if (tr && warp == 2 && lane == 0) ep.trace[3] = gtimer();
        run_epilogue<BN>(ep, tmem_base, tile_m, tile_n, M, N, tiles_a, warp & 3, (warp - 2) >> 2, lane);
        if (tr && warp == 2 && lane == 0) ep.trace[4] = gtimer();
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (tr && threadIdx.x == 0) ep.trace[5] = gtimer();  // nobody leaves (or frees TMEM) while the peer may still read its shared memory / signal its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor, `inner` contiguous elements per row, `outer` rows of pitch `ld` elements; box = {box_in, box_out}
int32_t make_map(CUtensorMap *map, const void *ptr, int64_t inner, int64_t outer, int64_t ld, int box_in, int box_out) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B200SURV_CUDA_ERROR; }
    B200_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
    B200_REQUIRE((ld * 2) % 16 == 0, "TMA operand row pitch must be a multiple of 8 bf16 elements");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_in, (cuuint32_t)box_out};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return B200SURV_CUDA_ERROR; }
    return B200SURV_OK;
}

template <bool A_MN, bool B_MN, int BN>
int32_t launch(const CUtensorMap &ma, const CUtensorMap &mb, int M, int N, int K, const Epilogue &ep, int splits,
               cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TileCfg<BN>::SMEM));
        attr_once.mark();
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    launch_chain(gemm_bf16_tc<A_MN, B_MN, BN>, grid, GEMM_THREADS, TileCfg<BN>::SMEM, st, ma, mb, M, N, K, ep);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}
template <int BN>
int32_t launch_bn(bool a_mn, bool b_mn, const CUtensorMap &ma, const CUtensorMap &mb, int M, int N, int K, const Epilogue &ep,
                  int splits, cudaStream_t st) {
    if (a_mn && b_mn) return launch<true, true, BN>(ma, mb, M, N, K, ep, splits, st);
    if (a_mn && !b_mn) return launch<true, false, BN>(ma, mb, M, N, K, ep, splits, st);
    if (!a_mn && b_mn) return launch<false, true, BN>(ma, mb, M, N, K, ep, splits, st);
    return launch<false, false, BN>(ma, mb, M, N, K, ep, splits, st);
}

}  // namespace

// C[M][N] = op(A) * op(B)^T-like contraction over K:
//   a_mn == 0: A is [M][K] (lda >= K)      a_mn == 1: A is [K][M] (lda >= M)
//   b_mn == 0: B is [N][K] (ldb >= K)      b_mn == 1: B is [K][N] (ldb >= N)
// slices a split-K launch of this shape uses (1 = not worth splitting): fill the SMs, at least 4 k-blocks a slice,
// no empty slice; *kb_per = k-blocks per slice
int splitk_slices(int M, int N, int K, int *kb_per) {
    constexpr int BN = 128;
    const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM), num_kb = (K + BK - 1) / BK;
    int s = num_sms() / tiles;
    if (s > num_kb / 4) s = num_kb / 4;
    if (s > 32) s = 32;
    if (s < 2) { if (kb_per) *kb_per = num_kb; return 1; }
    const int per = (num_kb + s - 1) / s;
    if (kb_per) *kb_per = per;
    return (num_kb + per - 1) / per;
}

template <bool A_MN, bool B_MN>
int32_t launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, int M, int N, int K, const Epilogue &ep, int splits,
                    cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
        attr_once.mark();
    }
    dim3 grid(2 * ((M + 255) / 256), (N + 255) / 256, splits);
    launch_chain(gemm_bf16_tc2<A_MN, B_MN>, grid, GEMM_THREADS, GEMM2_SMEM, st, ma, mb, M, N, K, ep);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

// tile_n: 128 (default), 192 or 256; 512 = CTA pairs (256 x 256 tiles, cta_group::2).  force_splits > 1: split K into exactly that many slices (caller sums the slices
// [force_splits][M][ldc] in `splitk_ws`); force_splits == 0 with splitk_ws != null: the heuristic of splitk_slices().
int32_t gemm_bf16_ex(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                     float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                     int tile_n, int force_splits, cudaStream_t st) {
    B200_REQUIRE(a && b && (c || c_bf16), "null pointer");
    B200_REQUIRE(M >= 1 && N >= 1 && K >= 1, "M, N, K must be positive");
    B200_REQUIRE(tile_n == 128 || tile_n == 192 || tile_n == 256 || tile_n == 512, "tile_n");
    const bool pair = tile_n == 512;
    CUtensorMap ma, mb;
    int32_t rc;
    rc = a_mn ? make_map(&ma, a, M, K, lda, 64, BK) : make_map(&ma, a, K, M, lda, BK, BM);
    if (rc) return rc;
    rc = b_mn ? make_map(&mb, b, N, K, ldb, 64, BK) : make_map(&mb, b, K, N, ldb, BK, pair ? 128 : tile_n);
    if (rc) return rc;
    Epilogue ep;
    ep.c = c; ep.c_bf16 = static_cast<__nv_bfloat16 *>(c_bf16); ep.bias = bias; ep.ldc = ldc; ep.ldc_bf16 = ldc_bf16;
    ep.relu = relu; ep.trace = g_gemm_trace;
    const int num_kb = (K + BK - 1) / BK;
    ep.kb_per = num_kb; ep.split_stride = 0;
    int splits = 1;
    if (splitk_ws != nullptr) {  // split-K into `splits` fp32 slices of [M][ldc] in the workspace (summed by the caller)
        B200_REQUIRE(c != nullptr && c_bf16 == nullptr && bias == nullptr && !relu, "split-K: plain fp32 output only");
        if (force_splits > 1) {
            ep.kb_per = (num_kb + force_splits - 1) / force_splits;
            splits = (num_kb + ep.kb_per - 1) / ep.kb_per;
            B200_REQUIRE(splits == force_splits, "force_splits leaves an empty K slice");
        } else {
            splits = splitk_slices(M, N, K, &ep.kb_per);
        }
        ep.c = splitk_ws; ep.split_stride = (int64_t)M * ldc;
    }
    const bool only_f32 = ep.c != nullptr && ep.c_bf16 == nullptr, only_bf16 = ep.c == nullptr && ep.c_bf16 != nullptr;
    ep.staged = ((only_f32 && (N & 3) == 0 && (ldc & 3) == 0 && (ep.split_stride & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(ep.c) & 15) == 0) ||
                 (only_bf16 && (N & 7) == 0 && (ldc_bf16 & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.c_bf16) & 15) == 0))
                    ? 1 : (only_f32 ? 2 : 0);
    if (pair) {
        if (a_mn && b_mn) return launch_pair<true, true>(ma, mb, M, N, K, ep, splits, st);
        if (a_mn && !b_mn) return launch_pair<true, false>(ma, mb, M, N, K, ep, splits, st);
        if (!a_mn && b_mn) return launch_pair<false, true>(ma, mb, M, N, K, ep, splits, st);
        return launch_pair<false, false>(ma, mb, M, N, K, ep, splits, st);
    }
    if (tile_n == 256) return launch_bn<256>(a_mn, b_mn, ma, mb, M, N, K, ep, splits, st);
    if (tile_n == 192) return launch_bn<192>(a_mn, b_mn, ma, mb, M, N, K, ep, splits, st);
    return launch_bn<128>(a_mn, b_mn, ma, mb, M, N, K, ep, splits, st);
}

int32_t gemm_bf16(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                  float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                  cudaStream_t st) {
    return gemm_bf16_ex(a, lda, a_mn, b, ldb, b_mn, M, N, K, c, ldc, c_bf16, ldc_bf16, bias, relu, splitk_ws, 128, 0, st);
}

}  // namespace b200surv

extern "C" int32_t b200surv_gemm_bf16(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb, int32_t b_mn,
                                      int32_t M, int32_t N, int32_t K, float *c, int64_t ldc, void *c_bf16,
                                      int64_t ldc_bf16, const float *bias, int32_t relu, b200surv_stream_t stream) {
    return b200surv::gemm_bf16(a, lda, a_mn, b, ldb, b_mn, M, N, K, c, ldc, c_bf16, ldc_bf16, bias, relu, nullptr,
                               b200surv::as_stream(stream));
}

/* diagnostics: CTA (0,0,0) of the pair kernel writes globaltimer stamps into `buf` (>= 32 int64 of device memory); null = off */
extern "C" void b200surv_debug_gemm_trace(long long *buf) { b200surv::g_gemm_trace = buf; }

extern "C" int32_t b200surv_gemm_bf16_ex(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb, int32_t b_mn,
                                         int32_t M, int32_t N, int32_t K, float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16,
                                         const float *bias, int32_t relu, int32_t tile_n, int32_t splits, float *slices,
                                         b200surv_stream_t stream) {
    if (splits >= 2) {
        if (slices == nullptr) { b200surv::set_error("b200surv_gemm_bf16_ex: splits >= 2 needs a slices buffer"); return B200SURV_BAD_ARG; }
        return b200surv::gemm_bf16_ex(a, lda, a_mn, b, ldb, b_mn, M, N, K, slices, ldc, nullptr, 0, nullptr, 0, slices, tile_n, splits,
                                      b200surv::as_stream(stream));
    }
    return b200surv::gemm_bf16_ex(a, lda, a_mn, b, ldb, b_mn, M, N, K, c, ldc, c_bf16, ldc_bf16, bias, relu, nullptr, tile_n, 0,
                                  b200surv::as_stream(stream));
}

// Split-K variant for outputs with few tiles and a long K (weight gradients): writes b200surv_gemm_splitk_slices(M, N, K)
// fp32 slices [M][ldc] back to back into `slices`; the caller sums them in slice order (deterministic).
extern "C" int32_t b200surv_gemm_splitk_slices(int32_t M, int32_t N, int32_t K) {
    if (M < 1 || N < 1 || K < 1) return 0;
    return b200surv::splitk_slices(M, N, K, nullptr);
}
extern "C" int32_t b200surv_gemm_bf16_splitk(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb,
                                             int32_t b_mn, int32_t M, int32_t N, int32_t K, float *slices, int64_t ldc,
                                             b200surv_stream_t stream) {
    return b200surv::gemm_bf16(a, lda, a_mn, b, ldb, b_mn, M, N, K, slices, ldc, nullptr, 0, nullptr, 0, slices,
                               b200surv::as_stream(stream));
}
