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
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier tx-count
//   warp 1      allocates 128 TMEM columns, issues tcgen05.mma.cta_group::1.kind::f16 (one elected
//               lane), tcgen05.commit frees smem stages / signals the accumulator
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias/ReLU -> global
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200surv {
namespace {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
constexpr int STAGES = 5;
constexpr int GEMM_THREADS = 192;
constexpr int TILE_BYTES = BM * BK * 2;  // 16 KB per operand per stage (BM == BN)
constexpr int TMEM_COLS = 128;
constexpr size_t GEMM_SMEM = (size_t)STAGES * 2 * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
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
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
};

// A_MN / B_MN: operand is stored [K][cols] (MN-major) instead of [rows][K] (K-major)
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N, int K,
             Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *tiles_a = smem;
    unsigned char *tiles_b = smem + (size_t)STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * 2 * TILE_BYTES);
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

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + STAGES + stage), phase ^ 1);
                const uint32_t full = smem_u32(bars + stage);
                mbar_expect_tx(full, 2 * TILE_BYTES);
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_BYTES);
                if (A_MN) {  // two [64 k][64 m] boxes
                    tma_load_2d(sa, &map_a, tile_m * BM, kb * BK, full);
                    tma_load_2d(sa + TILE_BYTES / 2, &map_a, tile_m * BM + 64, kb * BK, full);
                } else {     // one [128 m][64 k] box
                    tma_load_2d(sa, &map_a, kb * BK, tile_m * BM, full);
                }
                if (B_MN) {
                    tma_load_2d(sb, &map_b, tile_n * BN, kb * BK, full);
                    tma_load_2d(sb + TILE_BYTES / 2, &map_b, tile_n * BN + 64, kb * BK, full);
                } else {
                    tma_load_2d(sb, &map_b, kb * BK, tile_n * BN, full);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(bars + stage), phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(tiles_a + (size_t)stage * TILE_BYTES);
                const uint32_t sb = smem_u32(tiles_b + (size_t)stage * TILE_BYTES);
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
        const int q = warp & 3;
        mbar_wait(smem_u32(bars + 2 * STAGES), 0);
        tcgen05_fence_after();
        if (ep.staged) {
            // Coalesced stores: lane = row after tcgen05.ld, so a direct store touches 32 rows per instruction (32 L1 tag
            // cycles each; the wide dgrad outputs of the CT encoder were bound by exactly that).  Each warp stages its 32
            // rows x 128 columns in the pipeline's shared memory (free: every MMA has retired) and stores whole rows.
            const int ES = ep.c != nullptr ? 4 : 2, pitch = BN * ES + 16;
            unsigned char *stg = tiles_a + (size_t)q * (32 * (BN * 4 + 16));
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                const int col0 = tile_n * BN + c;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (ep.bias != nullptr && col0 + j < N) x += ep.bias[col0 + j];
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
            }
            __syncwarp();
            const int lanes_per_row = BN * ES / 16, per_vec = 16 / ES;      // 32 (fp32) or 16 (bf16) lanes store one row
            const int vec = lane % lanes_per_row, gcol = tile_n * BN + vec * per_vec;
            for (int k = lane / lanes_per_row; k < 32; k += 32 / lanes_per_row) {
                const int grow = tile_m * BM + q * 32 + k;
                if (grow < M && gcol < N) {                                 // N is a multiple of per_vec: whole vector inside
                    const uint4 u = *reinterpret_cast<const uint4 *>(stg + k * pitch + vec * 16);
                    unsigned char *dst = ES == 4
                        ? reinterpret_cast<unsigned char *>(ep.c + (int64_t)blockIdx.z * ep.split_stride + (int64_t)grow * ep.ldc + gcol)
                        : reinterpret_cast<unsigned char *>(ep.c_bf16 + (int64_t)grow * ep.ldc_bf16 + gcol);
                    *reinterpret_cast<uint4 *>(dst) = u;
                }
            }
        } else {
        const int row = tile_m * BM + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            const int col0 = tile_n * BN + c;
            if (row < M && col0 < N) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (ep.bias != nullptr && col0 + j < N) x += ep.bias[col0 + j];
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
        }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

template <bool A_MN, bool B_MN>
int32_t launch(const CUtensorMap &ma, const CUtensorMap &mb, int M, int N, int K, const Epilogue &ep, int splits,
               cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)GEMM_SMEM));
        attr_once.mark();
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    gemm_bf16_tc<A_MN, B_MN><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ma, mb, M, N, K, ep);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // namespace

// C[M][N] = op(A) * op(B)^T-like contraction over K:
//   a_mn == 0: A is [M][K] (lda >= K)      a_mn == 1: A is [K][M] (lda >= M)
//   b_mn == 0: B is [N][K] (ldb >= K)      b_mn == 1: B is [K][N] (ldb >= N)
// slices a split-K launch of this shape uses (1 = not worth splitting): fill the SMs, at least 4 k-blocks a slice,
// no empty slice; *kb_per = k-blocks per slice
int splitk_slices(int M, int N, int K, int *kb_per) {
    const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM), num_kb = (K + BK - 1) / BK;
    int s = num_sms() / tiles;
    if (s > num_kb / 4) s = num_kb / 4;
    if (s > 32) s = 32;
    if (s < 2) { if (kb_per) *kb_per = num_kb; return 1; }
    const int per = (num_kb + s - 1) / s;
    if (kb_per) *kb_per = per;
    return (num_kb + per - 1) / per;
}

int32_t gemm_bf16(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                  float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                  cudaStream_t st) {
    B200_REQUIRE(a && b && (c || c_bf16), "null pointer");
    B200_REQUIRE(M >= 1 && N >= 1 && K >= 1, "M, N, K must be positive");
    CUtensorMap ma, mb;
    int32_t rc;
    rc = a_mn ? make_map(&ma, a, M, K, lda, 64, BK) : make_map(&ma, a, K, M, lda, BK, BM);
    if (rc) return rc;
    rc = b_mn ? make_map(&mb, b, N, K, ldb, 64, BK) : make_map(&mb, b, K, N, ldb, BK, BN);
    if (rc) return rc;
    Epilogue ep;
    ep.c = c; ep.c_bf16 = static_cast<__nv_bfloat16 *>(c_bf16); ep.bias = bias; ep.ldc = ldc; ep.ldc_bf16 = ldc_bf16;
    ep.relu = relu;
    const int num_kb = (K + BK - 1) / BK;
    ep.kb_per = num_kb; ep.split_stride = 0;
    int splits = 1;
    if (splitk_ws != nullptr) {  // split-K into `splits` fp32 slices of [M][ldc] in the workspace (summed by the caller)
        B200_REQUIRE(c != nullptr && c_bf16 == nullptr && bias == nullptr && !relu, "split-K: plain fp32 output only");
        splits = splitk_slices(M, N, K, &ep.kb_per);
        ep.c = splitk_ws; ep.split_stride = (int64_t)M * ldc;
    }
    const bool only_f32 = ep.c != nullptr && ep.c_bf16 == nullptr, only_bf16 = ep.c == nullptr && ep.c_bf16 != nullptr;
    ep.staged = (only_f32 && (N & 3) == 0 && (ldc & 3) == 0 && (ep.split_stride & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(ep.c) & 15) == 0) ||
                (only_bf16 && (N & 7) == 0 && (ldc_bf16 & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.c_bf16) & 15) == 0);
    if (a_mn && b_mn) return launch<true, true>(ma, mb, M, N, K, ep, splits, st);
    if (a_mn && !b_mn) return launch<true, false>(ma, mb, M, N, K, ep, splits, st);
    if (!a_mn && b_mn) return launch<false, true>(ma, mb, M, N, K, ep, splits, st);
    return launch<false, false>(ma, mb, M, N, K, ep, splits, st);
}

}  // namespace b200surv

extern "C" int32_t b200surv_gemm_bf16(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb, int32_t b_mn,
                                      int32_t M, int32_t N, int32_t K, float *c, int64_t ldc, void *c_bf16,
                                      int64_t ldc_bf16, const float *bias, int32_t relu, b200surv_stream_t stream) {
    return b200surv::gemm_bf16(a, lda, a_mn, b, ldb, b_mn, M, N, K, c, ldc, c_bf16, ldc_bf16, bias, relu, nullptr,
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
