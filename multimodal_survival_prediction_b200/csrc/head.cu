// Gated late-fusion survival head, forward and backward, for B200.
//
// Replaces what PartialModalityNet.forward (scripts/training/partial_modality_training.py:234-277,
// layers :193-232) and MultiModalSurvivalNet.forward (scripts/training/final_multimodal.py:122-150)
// compute from the 128-d CT feature onward, plus the autograd backward of those layers:
//   rna 5005->512 (BatchNorm, ReLU, Dropout .3) ->128 (ReLU); clinical 1->32 (ReLU); per-modality mask
//   multiply; gate 291->64 (ReLU)->3 softmax over [ct, rna, clin, mask]; gate-weighted concat (288);
//   fusion 288->256 (BatchNorm, ReLU, Dropout .3) ->128 (ReLU); cox head 128->1.
// Every Linear layer and every weight/input gradient runs on the tensor cores through gemm_tc.cu
// (tcgen05 + TMEM + TMA, bf16 operands, fp32 accumulation); the kernels in this file are the fused
// element-wise / reduction glue between the GEMMs (BatchNorm statistics and application, dropout with a
// counter-based RNG that backward re-derives, mask/gate/softmax, the 128->1 Cox head).
//
// Launch structure (round 2; B = 4096: 16 launches forward, ~28 backward, was 21 + ~60):
//   * rna_encoder.0 forward and its weight gradient run on CTA PAIRS (256 x 256 tiles, tcgen05 cta_group::2); the
//     forward is split in two along K, and the kernel that forms the BatchNorm statistics also adds the two slices
//     and the bias (k_bn_stats_combine);
//   * every reduction over the batch (bias gradients, BatchNorm backward sums, the small gate / clinical / Cox-head
//     weight gradients) is fused into the kernel that produces the summand: a thread (or warp) owns a column and walks
//     a slice of rows, writes one fp64 partial per (row slice, column); one small kernel per group sums the slices in a
//     fixed order -- deterministic, no atomics;
//   * the four small weight-gradient GEMMs and the slice sums leave the critical path: they run on an internal side
//     stream forked from / joined to the caller's stream with events (also under CUDA-graph capture), while the
//     caller's stream carries the chain of input-gradient GEMMs.
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"

namespace b200surv {

int32_t gemm_bf16(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                  float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                  cudaStream_t st);
int32_t gemm_bf16_ex(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                     float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                     int tile_n, int force_splits, cudaStream_t st);
int splitk_slices(int M, int N, int K, int *kb_per);

namespace {

typedef __nv_bfloat16 bf16;
constexpr int H1 = 512, R1 = 128, CL = 32, CT = 128, FEAT = CT + R1 + CL;  // 288
constexpr int GZ = FEAT + 3, GZP = 296;                                      // gate input 291, padded to 296
constexpr int GH = 64, H2 = 256, F2N = 128;
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;
constexpr int RS_MAX = 128;  // row slices of the column reductions
constexpr int FS_MAX = 512;  // row slices of the fused backward reductions (8 rows per slice at B = 4096)
constexpr int SPLITK_ELEMS = 256 * 512;  // largest weight gradient that is split along K (all but rna_encoder.0)

__device__ __forceinline__ uint32_t rng32(uint64_t seed, uint32_t layer, uint64_t idx) {  // splitmix64
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1) + (uint64_t)layer * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}
__device__ __forceinline__ bool keep_elem(uint64_t seed, uint32_t layer, uint64_t idx, uint32_t thresh) {
    return rng32(seed, layer, idx) >= thresh;  // P(drop) = thresh / 2^32
}

// row = i / N without a 64-bit division where the flat index fits 32 bits (it does up to millions of rows)
__device__ __forceinline__ int64_t row_of(int64_t i, int N) {
    return i < 0x7fffffffll ? (int64_t)((unsigned)i / (unsigned)N) : i / N;
}

// ---------------------------------------------------------------- casts
// dst[r][0..Cp) = bf16(src[r][0..C)), zero padded; grid (column chunks, row groups): no per-element division
__global__ void k_cast_pad(const float *__restrict__ src, int64_t lds, bf16 *__restrict__ dst, int64_t ldd, int64_t R,
                           int C, int Cp) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    const int64_t step = gridDim.y;
    int64_t r = blockIdx.y;
    for (; r + 3 * step < R; r += 4 * step) {  // four rows in flight per thread
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = c < C ? src[(r + u * step) * lds + c] : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[(r + u * step) * ldd + c] = __float2bfloat16_rn(v[u]);
    }
    for (; r < R; r += step) dst[r * ldd + c] = __float2bfloat16_rn(c < C ? src[r * lds + c] : 0.f);
}
static void cast_pad(const float *src, int64_t lds, bf16 *dst, int64_t ldd, int64_t R, int C, int Cp, cudaStream_t st) {
    const unsigned gx = (unsigned)((Cp + 255) / 256);
    int64_t gy = R < 65535 ? R : 65535;
    const int64_t cap = (int64_t)32 * num_sms() / gx + 1;  // ~32 CTAs per SM are plenty
    if (gy > cap) gy = cap;
    launch_chain(k_cast_pad, dim3(gx, (unsigned)gy), 256, 0, st, src, lds, dst, ldd, R, C, Cp);
}

// A whole batch staged for a captured step in ONE launch: the RNA matrix as bf16 (4 columns per thread: scalar loads -- a
// row of 5005 floats starts on a 4-byte boundary only -- and one 8-byte store; two rows in flight) and up to three small
// fp32 inputs copied into the graph's static buffers by the blocks behind the cast blocks.
struct StageCopies { const float *src[3]; float *dst[3]; int64_t n[3]; int count; };
__global__ void __launch_bounds__(256)
k_stage_batch(const float *__restrict__ src, int64_t lds, bf16 *__restrict__ dst, int64_t ldd, int64_t R, int C, int Cp, int gx,
              int gy, const StageCopies cp) {
    pdl_prologue();
    const int ncast = gx * gy;
    if ((int)blockIdx.x >= ncast) {   // copy blocks
        const int64_t stride = (int64_t)(gridDim.x - ncast) * blockDim.x;
        for (int k = 0; k < cp.count; ++k)
            for (int64_t i = (int64_t)(blockIdx.x - ncast) * blockDim.x + threadIdx.x; i < cp.n[k]; i += stride) cp.dst[k][i] = cp.src[k][i];
        return;
    }
    const int bx = blockIdx.x % gx, by = blockIdx.x / gx;
    const int c = 4 * (bx * blockDim.x + threadIdx.x);
    if (c >= Cp) return;
    auto load4 = [&](int64_t r, float (&v)[4]) {
        const float *row = src + r * lds;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = c + u < C ? row[c + u] : 0.f;
    };
    auto store4 = [&](int64_t r, const float (&v)[4]) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&lo); pk.y = *reinterpret_cast<uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(dst + r * ldd + c) = pk;   // ldd and c are multiples of 4: 8-byte aligned
    };
    int64_t r = by;
    for (; r + gy < R; r += 2 * (int64_t)gy) {
        float a[4], b[4];
        load4(r, a); load4(r + gy, b);
        store4(r, a); store4(r + gy, b);
    }
    if (r < R) { float a[4]; load4(r, a); store4(r, a); }
}

// ---------------------------------------------------------------- column reductions (deterministic)
// partial[slice][2][N] (double).  MODE 0: (sum a, sum a^2)   MODE 1: (sum a, sum a * xhat), xhat from x, mu, rstd
// MODE 2: (sum a * s[row], sum s[row])   MODE 3: (sum a, 0)
template <int MODE>
__global__ void __launch_bounds__(256)
k_colreduce(const float *__restrict__ a, int64_t lda, const float *__restrict__ x, int64_t ldx,
            const float *__restrict__ mu, const float *__restrict__ rstd, const float *__restrict__ s, int64_t s_stride,
            int64_t B, int N, double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh0[8][32], sh1[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int slice = blockIdx.y, nslices = gridDim.y;
    const int64_t rows_per = (B + nslices - 1) / nslices;
    const int64_t r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    double v0 = 0.0, v1 = 0.0;
    if (n < N) {
        const float m = (MODE == 1) ? mu[n] : 0.f, rs = (MODE == 1) ? rstd[n] : 0.f;
        auto acc = [&](int64_t r, float av, float xv, float sv) {
            if (MODE == 0) { v0 += av; v1 += (double)av * av; }
            if (MODE == 1) { v0 += av; v1 += (double)av * ((xv - m) * rs); }
            if (MODE == 2) { v0 += (double)av * sv; v1 += sv; }
            if (MODE == 3) { v0 += av; }
        };
        int64_t r = r0 + ty;
        for (; r + 24 < r1; r += 32) {  // four rows in flight per thread
            float av[4], xv[4], sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                av[u] = a[(r + 8 * u) * lda + n];
                xv[u] = (MODE == 1) ? x[(r + 8 * u) * ldx + n] : 0.f;
                sv[u] = (MODE == 2) ? s[(r + 8 * u) * s_stride] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc(r + 8 * u, av[u], xv[u], sv[u]);
        }
        for (; r < r1; r += 8)
            acc(r, a[r * lda + n], (MODE == 1) ? x[r * ldx + n] : 0.f, (MODE == 2) ? s[r * s_stride] : 0.f);
    }
    sh0[ty][tx] = v0; sh1[ty][tx] = v1;
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { v0 += sh0[k][tx]; v1 += sh1[k][tx]; }
        partial[((size_t)slice * 2 + 0) * N + n] = v0;
        partial[((size_t)slice * 2 + 1) * N + n] = v1;
    }
}
// sums the slices: out0[n], out1[n] (float, nullable), scaled.  One WARP per column, lanes over the slices in a fixed
// order (deterministic); launched with colfinal_grid(N) blocks of 256 threads.
__device__ __forceinline__ void slice_sums(const double *__restrict__ partial, int nslices, int N, int n, int lane,
                                           double &v0, double &v1) {
    v0 = 0.0; v1 = 0.0;
    for (int k = lane; k < nslices; k += 32) { v0 += partial[((size_t)k * 2 + 0) * N + n]; v1 += partial[((size_t)k * 2 + 1) * N + n]; }
    v0 = warp_sum(v0); v1 = warp_sum(v1);
}
inline unsigned colfinal_grid(int N) { return (unsigned)((N + 7) / 8); }
__global__ void __launch_bounds__(256)
k_colreduce_final(const double *__restrict__ partial, int nslices, int N, float scale, float *out0, float *out1) {
    pdl_prologue();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    double v0, v1;
    slice_sums(partial, nslices, N, n, lane, v0, v1);
    if (lane == 0) {
        if (out0) out0[n] = (float)(v0 * scale);
        if (out1) out1[n] = (float)(v1 * scale);
    }
}

// ---------------------------------------------------------------- BatchNorm1d
// train: mu, rstd from the batch; running <- 0.9 running + 0.1 (mu, unbiased var).  eval: from running stats.
__global__ void __launch_bounds__(256)
k_bn_finalize(const double *__restrict__ partial, int nslices, int64_t B, int N, int training,
              float *__restrict__ run_mean, float *__restrict__ run_var, float *__restrict__ mu,
              float *__restrict__ rstd) {
    pdl_prologue();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    if (training) {
        double s, ss;
        slice_sums(partial, nslices, N, n, lane, s, ss);
        if (lane != 0) return;
        const double m = s / (double)B;
        double var = ss / (double)B - m * m;
        if (var < 0.0) var = 0.0;
        mu[n] = (float)m;
        rstd[n] = (float)(1.0 / sqrt(var + (double)BN_EPS));
        if (run_mean) {
            const double unb = var * ((double)B / (double)(B - 1));
            run_mean[n] = (1.f - BN_MOM) * run_mean[n] + BN_MOM * (float)m;
            run_var[n] = (1.f - BN_MOM) * run_var[n] + BN_MOM * (float)unb;
        }
    } else if (lane == 0) {
        mu[n] = run_mean[n];
        rstd[n] = rsqrtf(run_var[n] + BN_EPS);
    }
}
// y = dropout(relu(bn(x))) as bf16 (the next GEMM's A operand); optional keep-mask export for tests.
// Four consecutive columns per thread (N, ldx, ldy multiples of 4): 16-byte loads, 8-byte stores.
__global__ void k_bn_apply(const float *__restrict__ x, int64_t ldx, const float *__restrict__ mu,
                           const float *__restrict__ rstd, const float *__restrict__ gamma,
                           const float *__restrict__ beta, int64_t B, int N, uint32_t thresh, float inv_keep,
                           uint64_t seed, const uint64_t *__restrict__ seed_dev, uint32_t layer, bf16 *__restrict__ y,
                           int64_t ldy, uint8_t *__restrict__ keep_out) {
    pdl_prologue();
    if (seed_dev != nullptr) seed = *seed_dev;  // (CUDA-graph replays: the seed lives in device memory)
    const int n4 = N >> 2;
    const int64_t total = B * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_of(i, n4);
        const int n = (int)(i - r * n4) << 2;
        const float4 xv = *reinterpret_cast<const float4 *>(x + r * ldx + n);
        // (the per-column vectors are read element-wise: caller-owned parameter pointers need not be 16-byte aligned)
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = fmaxf((xs[u] - mu[n + u]) * rstd[n + u] * gamma[n + u] + beta[n + u], 0.f);
        uint8_t kp[4] = {1, 1, 1, 1};
        if (thresh) {
            const uint64_t e0 = (uint64_t)(r * N + n);  // element index r * N + column, as the backward kernels derive it
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool keep = keep_elem(seed, layer, e0 + u, thresh);
                kp[u] = keep ? 1 : 0;
                v[u] = keep ? v[u] * inv_keep : 0.f;
            }
        }
        if (keep_out) *reinterpret_cast<uchar4 *>(keep_out + r * N + n) = make_uchar4(kp[0], kp[1], kp[2], kp[3]);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        uint2 u2;
        u2.x = *reinterpret_cast<uint32_t *>(&p0); u2.y = *reinterpret_cast<uint32_t *>(&p1);
        *reinterpret_cast<uint2 *>(y + r * ldy + n) = u2;
    }
}
// ---------------------------------------------------------------- mask / clinical encoder / gate
// feat[b] = [ct*m0 (128) | R*m1 (128) | relu(clin*Wc+bc)*m2 (32)] (fp32); gated: z = bf16([feat | mask | 0 pad]);
// ungated (mask == nullptr): no masking, fused = bf16(feat)
__global__ void k_gate_prep(const float *__restrict__ ct, const float *__restrict__ R, const float *__restrict__ clin,
                            const float *__restrict__ mask, const float *__restrict__ wc, const float *__restrict__ bc,
                            int64_t B, float *__restrict__ feat, bf16 *__restrict__ z, bf16 *__restrict__ fused) {
    pdl_prologue();
    const int64_t total = B * GZP;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = row_of(i, GZP);
        const int j = (int)(i - b * GZP);
        float v = 0.f;
        if (j < FEAT) {
            if (j < CT) v = ct[b * CT + j] * (mask ? mask[b * 3 + 0] : 1.f);
            else if (j < CT + R1) v = R[b * R1 + (j - CT)] * (mask ? mask[b * 3 + 1] : 1.f);
            else {
                const int k = j - CT - R1;
                v = fmaxf(clin[b] * wc[k] + bc[k], 0.f) * (mask ? mask[b * 3 + 2] : 1.f);
            }
            feat[b * FEAT + j] = v;
            if (fused) fused[b * FEAT + j] = __float2bfloat16_rn(v);
        } else if (j < GZ) {
            v = mask ? mask[b * 3 + (j - FEAT)] : 0.f;
        }
        if (z) z[b * GZP + j] = __float2bfloat16_rn(v);
    }
}
// one warp per row: logits = zh Wg2^T + bg2, gate = softmax, fused = bf16(feat * gate[group])
__global__ void __launch_bounds__(256)
k_gate_apply(const float *__restrict__ zh, const float *__restrict__ wg2, const float *__restrict__ bg2,
             const float *__restrict__ feat, int64_t B, float *__restrict__ gate, float *__restrict__ gate_out,
             bf16 *__restrict__ fused) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < B; b += (int64_t)gridDim.x * 8) {
        const float z0 = zh[b * GH + lane], z1 = zh[b * GH + 32 + lane];
        float l0 = z0 * wg2[lane] + z1 * wg2[32 + lane];
        float l1 = z0 * wg2[GH + lane] + z1 * wg2[GH + 32 + lane];
        float l2 = z0 * wg2[2 * GH + lane] + z1 * wg2[2 * GH + 32 + lane];
        l0 = warp_sum(l0) + bg2[0]; l1 = warp_sum(l1) + bg2[1]; l2 = warp_sum(l2) + bg2[2];
        const float mx = fmaxf(l0, fmaxf(l1, l2));
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
        const float inv = 1.f / (e0 + e1 + e2);
        const float g0 = e0 * inv, g1 = e1 * inv, g2 = e2 * inv;
        if (lane == 0) {  // the saved copy (backward) and the caller's output
            gate[b * 3 + 0] = g0; gate[b * 3 + 1] = g1; gate[b * 3 + 2] = g2;
            gate_out[b * 3 + 0] = g0; gate_out[b * 3 + 1] = g1; gate_out[b * 3 + 2] = g2;
        }
        for (int j = lane; j < FEAT; j += 32) {
            const float g = j < CT ? g0 : (j < CT + R1 ? g1 : g2);
            fused[b * FEAT + j] = __float2bfloat16_rn(feat[b * FEAT + j] * g);
        }
    }
}
// ---------------------------------------------------------------- Cox head 128 -> 1
__global__ void __launch_bounds__(256)
k_cox_head(const float *__restrict__ f2, const float *__restrict__ w, const float *__restrict__ bias, int64_t B,
           float *__restrict__ hazard) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < B; b += (int64_t)gridDim.x * 8) {
        float s = 0.f;
        for (int j = lane; j < F2N; j += 32) s += f2[b * F2N + j] * w[j];
        s = warp_sum(s);
        if (lane == 0) hazard[b] = s + bias[0];
    }
}
// ---------------------------------------------------------------- BatchNorm1d, column-block kernels (round 2)
// A CTA owns BC = 128 consecutive columns: 32 column quads x 8 row lanes, 16-byte loads.  Two launches per direction
// instead of three: the kernel that CONSUMES the statistics (apply / dx) sums the row-slice partials of its own 128
// columns in its prologue (<= 32 slices, fixed order: deterministic) -- there is no separate finalize launch; the CTAs of
// the first row group also write what has to be kept (mu, rstd, running statistics / dbeta, dgamma, bias gradient).
constexpr int BC = 128, BS_MAX = 64, BR = 64;  // columns per CTA, row slices of the sums, rows per CTA of apply / dx
inline int bn_slices(int64_t B) { int64_t s = (B + 63) / 64; return (int)(s < 1 ? 1 : (s > BS_MAX ? BS_MAX : s)); }

// partial[slice][2][N] (fp64).  MODE 0: (sum a, sum a^2).  MODE 2: a = s0 + s1 + bias (the two K slices of the split
// rna_encoder.0 GEMM), stored to h, then as MODE 0 (want_stats == 0: only the combination).  MODE 1: (sum dy, sum dy * xhat)
// with dy = dA * keep/(1-p) * [bn(x) > 0] formed on the fly.  grid (N / 128, slices), 256 threads.
template <int MODE>
__global__ void __launch_bounds__(256)
k_bn_colstats(const float *__restrict__ a, const float *__restrict__ s1, const float *__restrict__ bias,
              const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ rstd,
              const float *__restrict__ gamma, const float *__restrict__ beta, int64_t B, int N, uint32_t thresh, float inv_keep,
              uint64_t seed, const uint64_t *__restrict__ seed_dev, uint32_t layer, int want_stats, float *__restrict__ h,
              double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh[8][2][BC];
    if (MODE == 1 && seed_dev != nullptr) seed = *seed_dev;
    const int tq = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * BC + 4 * tq, slice = blockIdx.y, nslices = gridDim.y;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    double v0[4] = {0.0, 0.0, 0.0, 0.0}, v1[4] = {0.0, 0.0, 0.0, 0.0};
    float cm[4] = {0.f, 0.f, 0.f, 0.f}, cr[4] = {0.f, 0.f, 0.f, 0.f}, cg[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (MODE == 2) cb[u] = bias[n + u];
        if (MODE == 1) { cm[u] = mu[n + u]; cr[u] = rstd[n + u]; cg[u] = gamma[n + u]; cb[u] = beta[n + u]; }
    }
    for (int64_t rb = r0 + ty; rb < r1; rb += 32) {  // four rows in flight per thread
        float4 av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t r = rb + 8 * q;
            av[q] = make_float4(0.f, 0.f, 0.f, 0.f); bv[q] = av[q];
            if (r < r1) {
                av[q] = *reinterpret_cast<const float4 *>(a + r * N + n);
                if (MODE == 2) bv[q] = *reinterpret_cast<const float4 *>(s1 + r * N + n);
                if (MODE == 1) bv[q] = *reinterpret_cast<const float4 *>(x + r * N + n);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t r = rb + 8 * q;
            if (r >= r1) continue;
            const float aa[4] = {av[q].x, av[q].y, av[q].z, av[q].w}, bb[4] = {bv[q].x, bv[q].y, bv[q].z, bv[q].w};
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (MODE == 0) { v0[u] += aa[u]; v1[u] += (double)aa[u] * aa[u]; }
                if (MODE == 2) { const float xx = (aa[u] + bb[u]) + cb[u]; o[u] = xx; v0[u] += xx; v1[u] += (double)xx * xx; }
                if (MODE == 1) {
                    const float xh = (bb[u] - cm[u]) * cr[u];
                    float g = aa[u];
                    if (thresh) g = keep_elem(seed, layer, (uint64_t)(r * N + n + u), thresh) ? g * inv_keep : 0.f;
                    const float dy = (xh * cg[u] + cb[u] > 0.f) ? g : 0.f;
                    v0[u] += dy; v1[u] += (double)dy * xh;
                }
            }
            if (MODE == 2) *reinterpret_cast<float4 *>(h + r * N + n) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    if (MODE == 2 && !want_stats) return;
#pragma unroll
    for (int u = 0; u < 4; ++u) { sh[ty][0][4 * tq + u] = v0[u]; sh[ty][1][4 * tq + u] = v1[u]; }
    __syncthreads();
    {
        const int c = threadIdx.x & (BC - 1), st = threadIdx.x >> 7;  // 256 threads = 128 columns x 2 sums
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += sh[k][st][c];
        partial[((size_t)slice * 2 + st) * N + blockIdx.x * BC + c] = v;
    }
}

// sums the row-slice partials of this CTA's 128 columns into shared memory: out[st][c], st = 0 / 1 (all threads call)
__device__ __forceinline__ void bn_block_sums(const double *__restrict__ partial, int nslices, int N, double (*out)[BC]) {
    const int c = threadIdx.x & (BC - 1), st = threadIdx.x >> 7;
    double pv[BS_MAX];  // all loads in flight together (one L2 round trip), then summed in slice order
#pragma unroll
    for (int k = 0; k < BS_MAX; ++k) pv[k] = k < nslices ? __ldcg(partial + ((size_t)k * 2 + st) * N + blockIdx.x * BC + c) : 0.0;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < BS_MAX; ++k) v += pv[k];
    out[st][c] = v;
    __syncthreads();
}

// y = dropout(relu(bn(x))) as bf16 (the next GEMM's A operand); optional keep-mask export for tests.
// train: mu, rstd from the partial sums of k_bn_colstats; running <- 0.9 running + 0.1 (mu, unbiased var).  eval: from the
// running statistics.  mu / rstd are kept for the backward pass.  grid (N / 128, ceil(B / 64)), 256 threads.
__global__ void __launch_bounds__(256)
k_bn_apply2(const float *__restrict__ x, const double *__restrict__ partial, int nslices, const float *__restrict__ gamma,
            const float *__restrict__ beta, int64_t B, int N, int training, float *__restrict__ run_mean,
            float *__restrict__ run_var, float *__restrict__ mu_out, float *__restrict__ rstd_out, uint32_t thresh,
            float inv_keep, uint64_t seed, const uint64_t *__restrict__ seed_dev, uint32_t layer, bf16 *__restrict__ y,
            uint8_t *__restrict__ keep_out) {
    pdl_prologue();
    __shared__ double sums[2][BC];
    __shared__ float s_mu[BC], s_rs[BC];
    if (seed_dev != nullptr) seed = *seed_dev;  // (CUDA-graph replays: the seed lives in device memory)
    const int cbase = blockIdx.x * BC;
    if (training) {
        bn_block_sums(partial, nslices, N, sums);
        if (threadIdx.x < BC) {
            const int c = threadIdx.x;
            const double m = sums[0][c] / (double)B;
            double var = sums[1][c] / (double)B - m * m;
            if (var < 0.0) var = 0.0;
            s_mu[c] = (float)m; s_rs[c] = (float)(1.0 / sqrt(var + (double)BN_EPS));
            if (blockIdx.y == 0) {
                mu_out[cbase + c] = s_mu[c]; rstd_out[cbase + c] = s_rs[c];
                if (run_mean) {
                    const double unb = var * ((double)B / (double)(B - 1));
                    run_mean[cbase + c] = (1.f - BN_MOM) * run_mean[cbase + c] + BN_MOM * (float)m;
                    run_var[cbase + c] = (1.f - BN_MOM) * run_var[cbase + c] + BN_MOM * (float)unb;
                }
            }
        }
    } else if (threadIdx.x < BC) {
        const int c = threadIdx.x;
        s_mu[c] = run_mean[cbase + c]; s_rs[c] = rsqrtf(run_var[cbase + c] + BN_EPS);
        if (blockIdx.y == 0) { mu_out[cbase + c] = s_mu[c]; rstd_out[cbase + c] = s_rs[c]; }
    }
    __syncthreads();
    const int tq = threadIdx.x & 31, ty = threadIdx.x >> 5, n = cbase + 4 * tq;
    float m4[4], r4[4], g4[4], b4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { m4[u] = s_mu[4 * tq + u]; r4[u] = s_rs[4 * tq + u]; g4[u] = gamma[n + u]; b4[u] = beta[n + u]; }
    const int64_t r0 = (int64_t)blockIdx.y * BR, r1 = min(B, r0 + BR);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + r * N + n);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float v[4];
        uint8_t kp[4] = {1, 1, 1, 1};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = fmaxf((xs[u] - m4[u]) * r4[u] * g4[u] + b4[u], 0.f);
            if (thresh) {
                const bool keep = keep_elem(seed, layer, (uint64_t)(r * N + n + u), thresh);
                kp[u] = keep ? 1 : 0;
                v[u] = keep ? v[u] * inv_keep : 0.f;
            }
        }
        if (keep_out) *reinterpret_cast<uchar4 *>(keep_out + r * N + n) = make_uchar4(kp[0], kp[1], kp[2], kp[3]);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        uint2 u2;
        u2.x = *reinterpret_cast<uint32_t *>(&p0); u2.y = *reinterpret_cast<uint32_t *>(&p1);
        *reinterpret_cast<uint2 *>(y + r * N + n) = u2;
    }
}

// BatchNorm backward, second launch: sdy = sum dy, sdyx = sum dy * xhat from the partials (prologue); the CTAs of the first
// row group write dbeta = sdy, dgamma = sdyx and the bias gradient of the Linear feeding the BatchNorm, sum_rows dx =
// (train ? 0 : gamma * rstd * sdy); then dx (bf16) = gamma * rstd * (dy - sdy / B - xhat * sdyx / B) (train) or
// gamma * rstd * dy (eval), dy re-derived.  grid (N / 128, ceil(B / 64)), 256 threads.
__global__ void __launch_bounds__(256)
k_bn_bwd_dx3(const float *__restrict__ dA, const float *__restrict__ x, const double *__restrict__ partial, int nslices,
             const float *__restrict__ mu, const float *__restrict__ rstd, const float *__restrict__ gamma,
             const float *__restrict__ beta, int64_t B, int N, int training, uint32_t thresh, float inv_keep, uint64_t seed,
             const uint64_t *__restrict__ seed_dev, uint32_t layer, float *__restrict__ dbeta, float *__restrict__ dgamma,
             float *__restrict__ dbias, bf16 *__restrict__ dx) {
    pdl_prologue();
    __shared__ double sums[2][BC];
    if (seed_dev != nullptr) seed = *seed_dev;
    const int cbase = blockIdx.x * BC;
    bn_block_sums(partial, nslices, N, sums);
    if (blockIdx.y == 0 && threadIdx.x < BC) {
        const int c = threadIdx.x;
        dbeta[cbase + c] = (float)sums[0][c]; dgamma[cbase + c] = (float)sums[1][c];
        dbias[cbase + c] = training ? 0.f : gamma[cbase + c] * rstd[cbase + c] * (float)sums[0][c];
    }
    const int tq = threadIdx.x & 31, ty = threadIdx.x >> 5, n = cbase + 4 * tq;
    const float invB = 1.f / (float)B;
    float m4[4], r4[4], g4[4], b4[4], sd[4], sx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        m4[u] = mu[n + u]; r4[u] = rstd[n + u]; g4[u] = gamma[n + u]; b4[u] = beta[n + u];
        sd[u] = (float)sums[0][4 * tq + u]; sx[u] = (float)sums[1][4 * tq + u];
    }
    const int64_t r0 = (int64_t)blockIdx.y * BR, r1 = min(B, r0 + BR);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float4 xv4 = *reinterpret_cast<const float4 *>(x + r * N + n), gv4 = *reinterpret_cast<const float4 *>(dA + r * N + n);
        const float xs[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, gs_[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - m4[u]) * r4[u];
            float g = gs_[u];
            if (thresh) g = keep_elem(seed, layer, (uint64_t)(r * N + n + u), thresh) ? g * inv_keep : 0.f;
            float v = (xh * g4[u] + b4[u] > 0.f) ? g : 0.f;
            if (training) v = v - sd[u] * invB - xh * sx[u] * invB;
            o[u] = v * g4[u] * r4[u];
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
        uint2 u2;
        u2.x = *reinterpret_cast<uint32_t *>(&p0); u2.y = *reinterpret_cast<uint32_t *>(&p1);
        *reinterpret_cast<uint2 *>(dx + r * N + n) = u2;
    }
}

// ---------------------------------------------------------------- fused reductions (round 2)
// one launch for the bf16 copies of all weight matrices (K padded with zeros where the TMA row pitch needs it)
// (segment k owns the blocks [b0[k], b0[k + 1]): all matrices are converted at the same time)
struct CastSeg { const float *src; bf16 *dst; int lds, ldd, R, C, Cp; };
// A block converts CAST_ROWS rows x 256 columns of its matrix: thread = column, the rows' loads in flight together.
constexpr int CAST_ROWS = 8;
struct CastSegs { CastSeg s[5]; int b0[6]; int n; };
__global__ void __launch_bounds__(256)
k_cast_multi(const CastSegs segs, uint64_t *seed_advance) {
    pdl_prologue();
    if (seed_advance != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *seed_advance += 1;  // (see B200SURV_HEAD_SEED_ADVANCE)
    int k = 0;
#pragma unroll
    for (int q = 1; q < 5; ++q) k += (q < segs.n && (int)blockIdx.x >= segs.b0[q]) ? 1 : 0;
    const CastSeg g = segs.s[k];
    const int bx = blockIdx.x - segs.b0[k], chunks = (g.Cp + 255) / 256;
    const int rg = bx / chunks, c = (bx - rg * chunks) * 256 + threadIdx.x;
    if (c >= g.Cp) return;
    const int r0 = rg * CAST_ROWS;
    float v[CAST_ROWS];
#pragma unroll
    for (int u = 0; u < CAST_ROWS; ++u) v[u] = (r0 + u < g.R && c < g.C) ? g.src[(int64_t)(r0 + u) * g.lds + c] : 0.f;
#pragma unroll
    for (int u = 0; u < CAST_ROWS; ++u)
        if (r0 + u < g.R) g.dst[(int64_t)(r0 + u) * g.ldd + c] = __float2bfloat16_rn(v[u]);
}

// h[r][n] = s0[r][n] + s1[r][n] + bias[n] (the two K slices of the split GEMM), and the BatchNorm partial sums
// (sum h, sum h^2) per (row slice, column) in the layout of k_colreduce<0>.  grid (N / 32, nslices), 256 threads.
__global__ void __launch_bounds__(256)
k_bn_stats_combine(const float *__restrict__ s0, const float *__restrict__ s1, const float *__restrict__ bias, int64_t B, int N,
                   int want_stats, float *__restrict__ h, double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh0[8][32], sh1[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx, slice = blockIdx.y, nslices = gridDim.y;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    double v0 = 0.0, v1 = 0.0;
    if (n < N) {
        const float bn = bias[n];
        int64_t r = r0 + ty;
        for (; r + 24 < r1; r += 32) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = s0[(r + 8 * u) * N + n]; b[u] = s1[(r + 8 * u) * N + n]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float x = (a[u] + b[u]) + bn;
                h[(r + 8 * u) * N + n] = x;
                v0 += x; v1 += (double)x * x;
            }
        }
        for (; r < r1; r += 8) {
            const float x = (s0[r * N + n] + s1[r * N + n]) + bn;
            h[r * N + n] = x;
            v0 += x; v1 += (double)x * x;
        }
    }
    if (!want_stats) return;
    sh0[ty][tx] = v0; sh1[ty][tx] = v1;
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { v0 += sh0[k][tx]; v1 += sh1[k][tx]; }
        partial[((size_t)slice * 2 + 0) * N + n] = v0;
        partial[((size_t)slice * 2 + 1) * N + n] = v1;
    }
}

// BatchNorm backward, pass 1: dy = dA * keep/(1-p) * [bn(x) > 0] formed on the fly (not stored), partial sums
// (sum dy, sum dy * xhat) per (row slice, column).  grid (N / 32, nslices), 256 threads.
__global__ void __launch_bounds__(256)
k_bn_bwd_stats(const float *__restrict__ dA, int64_t ldd, const float *__restrict__ x, int64_t ldx, const float *__restrict__ mu,
               const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ beta, int64_t B, int N,
               uint32_t thresh, float inv_keep, uint64_t seed, const uint64_t *__restrict__ seed_dev, uint32_t layer,
               double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh0[8][32], sh1[8][32];
    if (seed_dev != nullptr) seed = *seed_dev;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx, slice = blockIdx.y, nslices = gridDim.y;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    double v0 = 0.0, v1 = 0.0;
    if (n < N) {
        const float m = mu[n], rs = rstd[n], ga = gamma[n], be = beta[n];
        for (int64_t rb = r0 + ty; rb < r1; rb += 32) {  // four rows in flight per thread
            float xv[4], gv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = rb + 8 * u;
                xv[u] = r < r1 ? x[r * ldx + n] : 0.f; gv[u] = r < r1 ? dA[r * ldd + n] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = rb + 8 * u;
                if (r < r1) {
                    const float xh = (xv[u] - m) * rs;
                    float g = gv[u];
                    if (thresh) g = keep_elem(seed, layer, (uint64_t)(r * N + n), thresh) ? g * inv_keep : 0.f;
                    const float dy = (xh * ga + be > 0.f) ? g : 0.f;
                    v0 += dy; v1 += (double)dy * xh;
                }
            }
        }
    }
    sh0[ty][tx] = v0; sh1[ty][tx] = v1;
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { v0 += sh0[k][tx]; v1 += sh1[k][tx]; }
        partial[((size_t)slice * 2 + 0) * N + n] = v0;
        partial[((size_t)slice * 2 + 1) * N + n] = v1;
    }
}
// pass 2: dbeta = sum dy, dgamma = sum dy * xhat (one warp per column, slices in a fixed order) and the bias gradient
// of the Linear that feeds the BatchNorm: sum_rows dx = (train ? 0 : gamma * rstd * sum dy)
__global__ void __launch_bounds__(256)
k_bn_bwd_final(const double *__restrict__ partial, int nslices, int N, const float *__restrict__ gamma,
               const float *__restrict__ rstd, int training, float *__restrict__ dbeta, float *__restrict__ dgamma,
               float *__restrict__ dbias) {
    pdl_prologue();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    double v0, v1;
    slice_sums(partial, nslices, N, n, lane, v0, v1);
    if (lane == 0) {
        dbeta[n] = (float)v0; dgamma[n] = (float)v1;
        dbias[n] = training ? 0.f : gamma[n] * rstd[n] * (float)v0;
    }
}
// pass 3: dx (bf16) = gamma * rstd * (dy - sdy / B - xhat * sdyx / B) (train) or gamma * rstd * dy (eval), dy re-derived.
// Four consecutive columns per thread (N and the pitches are multiples of 4).
__global__ void k_bn_bwd_dx2(const float *__restrict__ dA, int64_t ldd, const float *__restrict__ x, int64_t ldx,
                             const float *__restrict__ mu, const float *__restrict__ rstd, const float *__restrict__ gamma,
                             const float *__restrict__ beta, const float *__restrict__ sdy, const float *__restrict__ sdyx,
                             int64_t B, int N, int training, uint32_t thresh, float inv_keep, uint64_t seed,
                             const uint64_t *__restrict__ seed_dev, uint32_t layer, bf16 *__restrict__ dx, int64_t lddx) {
    pdl_prologue();
    if (seed_dev != nullptr) seed = *seed_dev;
    const int n4 = N >> 2;
    const int64_t total = B * n4;
    const float invB = 1.f / (float)B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_of(i, n4);
        const int n = (int)(i - r * n4) << 2;
        const float4 xv4 = *reinterpret_cast<const float4 *>(x + r * ldx + n), g4 = *reinterpret_cast<const float4 *>(dA + r * ldd + n);
        // (the per-column vectors are read element-wise: the gradient vectors sit at arbitrary offsets of the caller's buffer)
        const float xs[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, gs_[4] = {g4.x, g4.y, g4.z, g4.w};
        float ms[4], rs[4], ga[4], be[4], sd[4], sx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ms[u] = mu[n + u]; rs[u] = rstd[n + u]; ga[u] = gamma[n + u]; be[u] = beta[n + u];
            sd[u] = sdy[n + u]; sx[u] = sdyx[n + u];
        }
        const uint64_t e0 = (uint64_t)(r * N + n);
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - ms[u]) * rs[u];
            float g = gs_[u];
            if (thresh) g = keep_elem(seed, layer, e0 + u, thresh) ? g * inv_keep : 0.f;
            float v = (xh * ga[u] + be[u] > 0.f) ? g : 0.f;
            if (training) v = v - sd[u] * invB - xh * sx[u] * invB;
            o[u] = v * ga[u] * rs[u];
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
        uint2 u2;
        u2.x = *reinterpret_cast<uint32_t *>(&p0); u2.y = *reinterpret_cast<uint32_t *>(&p1);
        *reinterpret_cast<uint2 *>(dx + r * lddx + n) = u2;
    }
}

// Cox head backward, one launch: df2 = bf16(dhz[b] * w[j] * [f2 > 0]) and the partial sums per (row slice, column j) of
//   [0] sum_b dhz[b] * f2[b][j]  (cox_head.weight)   [1] sum_b df2[b][j]  (fusion.4.bias: column sums of the SAME bf16
//   values the GEMMs read)   [2][0] sum_b dhz[b]  (cox_head.bias).  grid (nslices), 128 threads: a thread owns column j.
__global__ void __launch_bounds__(F2N)
k_cox_bwd_fused(const float *__restrict__ dhz, const float *__restrict__ f2, const float *__restrict__ w, int64_t B,
                bf16 *__restrict__ df2, double *__restrict__ partial) {
    pdl_prologue();
    const int j = threadIdx.x, slice = blockIdx.x, nslices = gridDim.x;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    const float wj = w[j];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int64_t b0 = r0; b0 < r1; b0 += 8) {  // the loads of eight rows in flight together
        float d[8], f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = b0 + u < r1;
            d[u] = in ? dhz[b0 + u] : 0.f; f[u] = in ? f2[(b0 + u) * F2N + j] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (b0 + u < r1) {
                const bf16 q = __float2bfloat16_rn(f[u] > 0.f ? d[u] * wj : 0.f);
                df2[(b0 + u) * F2N + j] = q;
                a0 += (double)d[u] * f[u]; a1 += (double)__bfloat162float(q); a2 += d[u];
            }
        }
    }
    double *p = partial + (size_t)slice * (3 * F2N);
    p[j] = a0; p[F2N + j] = a1;
    if (j == 0) p[2 * F2N] = a2;
}

// sums the row slices of `ncols` columns (partial[slice][ncols], fp64) into up to four float outputs: segment s takes the
// columns [c0[s], c0[s + 1]).  One warp per column, slices in a fixed order.
struct SumSegs { float *out[4]; int c0[5]; int n; };
__global__ void __launch_bounds__(256)
k_sum_slices(const double *__restrict__ partial, int nslices, int ncols, const SumSegs segs) {
    pdl_prologue();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= ncols) return;
    double v = 0.0;
    for (int k = lane; k < nslices; k += 32) v += partial[(size_t)k * ncols + c];
    v = warp_sum(v);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < segs.n && c >= segs.c0[s] && c < segs.c0[s + 1]) segs.out[s][c - segs.c0[s]] = (float)v;
    }
}

// Gate backward (softmax + scaling), one launch, one warp per row inside a row slice per CTA:
//   dfeat = dfused * gate (fp32 [B][288]); dzh = bf16((dlogit Wg2) * [zh > 0]) [B][64]; partial sums per row slice of
//   gate.2.weight [3][64] (sum_b dlogit[b][k] zh[b][c]), gate.2.bias [3], gate.0.bias [64] (column sums of dzh as bf16).
// partial[slice][GP_COLS]: [0,192) gate2_w, [192,195) gate2_b, [195,259) gate0_b.  grid (nslices), 256 threads.
constexpr int GP_COLS = 3 * GH + 3 + GH;
__global__ void __launch_bounds__(256)
k_gate_bwd_fused(const float *__restrict__ dfused, const float *__restrict__ feat, const float *__restrict__ gate,
                 const float *__restrict__ zh, const float *__restrict__ wg2, const float *__restrict__ d_gate_ext, int64_t B,
                 float *__restrict__ dfeat, bf16 *__restrict__ dzh, double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh[8][GP_COLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slice = blockIdx.x, nslices = gridDim.x;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    double aw[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}, ab[3] = {0.0, 0.0, 0.0}, az[2] = {0.0, 0.0};
    for (int64_t b = r0 + warp; b < r1; b += 8) {
        const float g0 = gate[b * 3], g1 = gate[b * 3 + 1], g2 = gate[b * 3 + 2];
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        for (int j = lane; j < FEAT; j += 32) {
            const float df = dfused[b * FEAT + j], f = feat[b * FEAT + j];
            if (j < CT) { d0 += df * f; dfeat[b * FEAT + j] = df * g0; }
            else if (j < CT + R1) { d1 += df * f; dfeat[b * FEAT + j] = df * g1; }
            else { d2 += df * f; dfeat[b * FEAT + j] = df * g2; }
        }
        d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
        if (d_gate_ext) { d0 += d_gate_ext[b * 3]; d1 += d_gate_ext[b * 3 + 1]; d2 += d_gate_ext[b * 3 + 2]; }
        const float dot = g0 * d0 + g1 * d1 + g2 * d2;
        const float dl[3] = {g0 * (d0 - dot), g1 * (d1 - dot), g2 * (d2 - dot)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            const float z = zh[b * GH + k];
            const float v = dl[0] * wg2[k] + dl[1] * wg2[GH + k] + dl[2] * wg2[2 * GH + k];
            const bf16 q = __float2bfloat16_rn(z > 0.f ? v : 0.f);
            dzh[b * GH + k] = q;
            az[u] += (double)__bfloat162float(q);
#pragma unroll
            for (int c = 0; c < 3; ++c) aw[c][u] += (double)dl[c] * z;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) ab[c] += dl[c];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int c = 0; c < 3; ++c) sh[warp][c * GH + lane + 32 * u] = aw[c][u];
        sh[warp][3 * GH + 3 + lane + 32 * u] = az[u];
    }
    if (lane < 3) sh[warp][3 * GH + lane] = ab[lane];
    __syncthreads();
    for (int c = threadIdx.x; c < GP_COLS; c += blockDim.x) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += sh[k][c];
        partial[(size_t)slice * GP_COLS + c] = v;
    }
}

// Mask / clinical-encoder backward, one launch: dtot = dfeat (+ dz[:, :288]); d_ct = dtot[0:128] * m0;
// dR = bf16(dtot[128:256] * m1 * [R > 0]); dC = dtot[256:288] * m2 * [C > 0] (not stored), and the partial sums per row slice
// of rna_encoder.4.bias (column sums of dR as bf16) [128], clinical_encoder.0.weight [32] (sum_b dC clin[b]) and .bias [32].
// partial[slice][PP_COLS].  grid (nslices), 288 threads: a thread owns column j of the 288.
constexpr int PP_COLS = R1 + 2 * CL;
__global__ void __launch_bounds__(FEAT)
k_prep_bwd_fused(const float *__restrict__ dfeat, const float *__restrict__ dz, int64_t lddz, const float *__restrict__ mask,
                 const float *__restrict__ R, const float *__restrict__ clin, const float *__restrict__ wc,
                 const float *__restrict__ bc, int64_t B, float *__restrict__ d_ct, bf16 *__restrict__ dR,
                 double *__restrict__ partial) {
    pdl_prologue();
    const int j = threadIdx.x, slice = blockIdx.x, nslices = gridDim.x;
    const int64_t rows_per = (B + nslices - 1) / nslices, r0 = slice * rows_per, r1 = min(B, r0 + rows_per);
    const int grp = j < CT ? 0 : (j < CT + R1 ? 1 : 2);
    const int k = grp == 0 ? j : (grp == 1 ? j - CT : j - CT - R1);
    const float wck = grp == 2 ? wc[k] : 0.f, bck = grp == 2 ? bc[k] : 0.f;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t b0 = r0; b0 < r1; b0 += 8) {  // the loads of eight rows in flight together
        float v[8], aux[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t b = b0 + u;
            const bool in = b < r1;
            v[u] = in ? dfeat[b * FEAT + j] + (dz ? dz[b * lddz + j] : 0.f) : 0.f;
            v[u] *= (in && mask) ? mask[b * 3 + grp] : 1.f;
            aux[u] = !in ? 0.f : (grp == 1 ? R[b * R1 + k] : (grp == 2 ? clin[b] : 0.f));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t b = b0 + u;
            if (b < r1) {
                if (grp == 0) {
                    if (d_ct) d_ct[b * CT + k] = v[u];
                } else if (grp == 1) {
                    const bf16 q = __float2bfloat16_rn(aux[u] > 0.f ? v[u] : 0.f);
                    dR[b * R1 + k] = q;
                    a0 += (double)__bfloat162float(q);
                } else {
                    const float dc = (aux[u] * wck + bck > 0.f) ? v[u] : 0.f;
                    a0 += (double)dc * aux[u]; a1 += dc;
                }
            }
        }
    }
    double *p = partial + (size_t)slice * PP_COLS;
    if (grp == 1) p[k] = a0;
    if (grp == 2) { p[R1 + k] = a0; p[R1 + CL + k] = a1; }
}

// ---------------------------------------------------------------- host orchestration
inline int gs(int64_t total) {  // grid for grid-stride element-wise kernels
    int64_t g = (total + 255) / 256;
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline int nslices_for(int64_t B) { int64_t s = (B + 31) / 32; return (int)(s < 1 ? 1 : (s > RS_MAX ? RS_MAX : s)); }
inline int fslices_for(int64_t B) { int64_t s = (B + 7) / 8; return (int)(s < 1 ? 1 : (s > FS_MAX ? FS_MAX : s)); }

struct Carver {
    unsigned char *base;
    size_t off;
    template <typename T>
    T *take(size_t count) {
        T *p = reinterpret_cast<T *>(base + off);
        off = align_up(off + count * sizeof(T), 256);
        return p;
    }
};

// activations kept from forward to backward (the `saved` buffer)
struct Saved {
    bf16 *xb, *w1b, *w2b, *wg1b, *wf1b, *wf2b;  // bf16 operand copies: x [B][Kp], weights
    float *h1, *mu1, *rstd1;                     // pre-BN activations and statistics
    bf16 *a1;                                    // [B][512]
    float *r;                                    // [B][128] relu(rna_encoder.4)
    float *feat;                                 // [B][288]
    bf16 *z;                                     // [B][296]
    float *zh, *gate;                            // [B][64], [B][3]
    bf16 *fused;                                 // [B][288]
    float *h2, *mu2, *rstd2;
    bf16 *a2;                                    // [B][256]
    float *f2;                                   // [B][128]
};
inline int kpad(int rna_dim) { return (rna_dim + 7) / 8 * 8; }
Saved carve_saved(void *buf, int64_t B, int rna_dim, size_t *bytes) {
    Carver c{static_cast<unsigned char *>(buf), 0};
    const int Kp = kpad(rna_dim);
    Saved s;
    s.xb = c.take<bf16>((size_t)B * Kp); s.w1b = c.take<bf16>((size_t)H1 * Kp); s.w2b = c.take<bf16>((size_t)R1 * H1);
    s.wg1b = c.take<bf16>((size_t)GH * GZP); s.wf1b = c.take<bf16>((size_t)H2 * FEAT); s.wf2b = c.take<bf16>((size_t)F2N * H2);
    s.h1 = c.take<float>((size_t)B * H1); s.mu1 = c.take<float>(H1); s.rstd1 = c.take<float>(H1);
    s.a1 = c.take<bf16>((size_t)B * H1); s.r = c.take<float>((size_t)B * R1); s.feat = c.take<float>((size_t)B * FEAT);
    s.z = c.take<bf16>((size_t)B * GZP); s.zh = c.take<float>((size_t)B * GH); s.gate = c.take<float>((size_t)B * 3);
    s.fused = c.take<bf16>((size_t)B * FEAT); s.h2 = c.take<float>((size_t)B * H2); s.mu2 = c.take<float>(H2);
    s.rstd2 = c.take<float>(H2); s.a2 = c.take<bf16>((size_t)B * H2); s.f2 = c.take<float>((size_t)B * F2N);
    *bytes = c.off;
    return s;
}

// scratch used inside one call (the `workspace` buffer)
struct Scratch {
    double *partial;                  // [RS_MAX][2][512]: BatchNorm sums (caller's stream)
    double *p_cox, *p_gate, *p_prep;  // row-slice partials of the fused reductions (summed on the side stream)
    float *t0, *t1;                   // [B][512] fp32 temporaries (caller's stream)
    float *slices;                    // [2][B][512]: the two K slices of rna_encoder.0 forward
    bf16 *df2, *dh2, *dzh, *dR, *dh1; // dY operands of the gradient GEMMs: one buffer each (the side stream reads them)
    float *splitk;                    // [32 slices][SPLITK_ELEMS] fp32 partial weight gradients (side stream)
};
Scratch carve_scratch(void *buf, int64_t B, size_t *bytes) {
    Carver c{static_cast<unsigned char *>(buf), 0};
    Scratch s;
    s.partial = c.take<double>((size_t)RS_MAX * 2 * H1);
    s.p_cox = c.take<double>((size_t)FS_MAX * 3 * F2N);
    s.p_gate = c.take<double>((size_t)FS_MAX * GP_COLS);
    s.p_prep = c.take<double>((size_t)FS_MAX * PP_COLS);
    s.t0 = c.take<float>((size_t)B * H1); s.t1 = c.take<float>((size_t)B * H1);
    s.slices = c.take<float>((size_t)2 * B * H1);
    s.df2 = c.take<bf16>((size_t)B * F2N); s.dh2 = c.take<bf16>((size_t)B * H2); s.dzh = c.take<bf16>((size_t)B * GH);
    s.dR = c.take<bf16>((size_t)B * R1); s.dh1 = c.take<bf16>((size_t)B * H1);
    s.splitk = c.take<float>((size_t)32 * SPLITK_ELEMS);
    *bytes = c.off;
    return s;
}

// the side stream of the backward pass (small weight-gradient GEMMs, slice sums), created once per device
struct HeadLanes {
    cudaStream_t side;
    cudaEvent_t fork, join;
};
HeadLanes *head_lanes() {
    static HeadLanes pool[MAX_DEVICES];
    static int state[MAX_DEVICES];  // 0 = not tried, 1 = ready, -1 = failed
    static std::mutex mu;
    const int dev = current_device();
    if (dev < 0 || dev >= MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (state[dev] == 0) {
        const bool ok = cudaStreamCreateWithFlags(&pool[dev].side, cudaStreamNonBlocking) == cudaSuccess &&
                        cudaEventCreateWithFlags(&pool[dev].fork, cudaEventDisableTiming) == cudaSuccess &&
                        cudaEventCreateWithFlags(&pool[dev].join, cudaEventDisableTiming) == cudaSuccess;
        state[dev] = ok ? 1 : -1;
    }
    return state[dev] == 1 ? &pool[dev] : nullptr;
}

template <int MODE>
void colreduce(const float *a, int64_t lda, const float *x, int64_t ldx, const float *mu, const float *rstd,
               const float *s, int64_t s_stride, int64_t B, int N, double *partial, int nsl, cudaStream_t st) {
    launch_chain(k_colreduce<MODE>, dim3((N + 31) / 32, nsl), 256, 0, st, a, lda, x, ldx, mu, rstd, s, s_stride, B, N, partial);
}
// weight gradient dW[M][N] = A^T B over K = batch rows (both operands MN-major).  The small ones (one to six output
// tiles for 64 k-blocks) are split along K over the SMs into fp32 slices and summed in a fixed order (deterministic).
__global__ void k_splitk_sum(const float *__restrict__ part, int slices, int64_t elems, float *__restrict__ out) {
    pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int k = 0; k < slices; ++k) v += part[(size_t)k * elems + i];
        out[i] = v;
    }
}
int32_t gemm_wgrad(const void *a, int64_t lda, const void *b, int64_t ldb, int M, int N, int K, float *c, int64_t ldc,
                   float *splitk_ws, cudaStream_t st) {
    const int slices = ((int64_t)M * ldc <= SPLITK_ELEMS) ? splitk_slices(M, N, K, nullptr) : 1;
    if (slices == 1) return gemm_bf16(a, lda, 1, b, ldb, 1, M, N, K, c, ldc, nullptr, 0, nullptr, 0, nullptr, st);
    const int32_t rc = gemm_bf16(a, lda, 1, b, ldb, 1, M, N, K, c, ldc, nullptr, 0, nullptr, 0, splitk_ws, st);
    if (rc) return rc;
    const int64_t elems = (int64_t)M * ldc;
    launch_chain(k_splitk_sum, (unsigned)((elems + 255) / 256), 256, 0, st, splitk_ws, slices, elems, c);
    return B200SURV_OK;
}

// ---------------------------------------------------------------- gate-entropy regulariser
// partial_modality_training.py:322-331: loss = -mean_b( -sum_k g log(g + eps) ) = mean_b sum_k g log(g + eps).
// One CTA (the value is one scalar; 3 B elements), fixed summation order: deterministic.
__global__ void __launch_bounds__(1024)
k_gate_entropy_fwd(const float *__restrict__ gate, int64_t B, float eps, float *__restrict__ out) {
    pdl_prologue();
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < 3 * B; i += blockDim.x) {
        const float g = gate[i];
        s += (double)(g * logf(g + eps));
    }
    s = block_reduce<double>(s, 0.0, OpAddD(), red);
    if (threadIdx.x == 0) out[0] = (float)(s / (double)B);
}
// d loss / d g = (log(g + eps) + g / (g + eps)) / B
__global__ void k_gate_entropy_bwd(const float *__restrict__ gate, const float *__restrict__ grad_out, int64_t B, float eps,
                                   float *__restrict__ d_gate) {
    pdl_prologue();
    const float k = grad_out[0] / (float)B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * B; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = gate[i], ge = g + eps;
        d_gate[i] = k * (logf(ge) + g / ge);
    }
}

uint32_t drop_thresh(float p) {
    if (p <= 0.f) return 0;
    double t = (double)p * 4294967296.0;
    if (t > 4294967295.0) t = 4294967295.0;
    return (uint32_t)t;
}

}  // namespace

}  // namespace b200surv

using namespace b200surv;

extern "C" {

size_t b200surv_head_saved_bytes(int64_t B, int32_t rna_dim) {
    size_t bytes = 0;
    carve_saved(nullptr, B, rna_dim, &bytes);
    return bytes;
}
size_t b200surv_head_workspace_bytes(int64_t B, int32_t rna_dim) {
    (void)rna_dim;
    size_t bytes = 0;
    carve_scratch(nullptr, B, &bytes);
    return bytes;
}

int32_t b200surv_head_fwd(const b200surv_head_params *p, const float *ct_feat, const float *rna, const float *clinical,
                          const float *mask, int64_t B, int32_t rna_dim, int32_t training, float dropout_p,
                          uint64_t seed, float *hazard, float *gate, uint8_t *keep1, uint8_t *keep2, void *saved,
                          size_t saved_bytes, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    const bool x_staged = (training & B200SURV_HEAD_X_STAGED) != 0, seed_adv = (training & B200SURV_HEAD_SEED_ADVANCE) != 0;
    training &= ~(B200SURV_HEAD_X_STAGED | B200SURV_HEAD_SEED_ADVANCE);
    B200_REQUIRE(p && ct_feat && (rna || x_staged) && clinical && hazard && saved && workspace, "null pointer");
    B200_REQUIRE(B >= 1 && B < (int64_t)1 << 24, "batch size must be in [1, 2^24)");
    B200_REQUIRE(rna_dim >= 1 && rna_dim <= 65536, "rna_dim");
    B200_REQUIRE(!(training && B < 2), "Expected more than 1 value per channel when training (BatchNorm1d)");
    B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p");
    const bool gated = mask != nullptr;
    B200_REQUIRE(!gated || (p->gate0_w && p->gate0_b && p->gate2_w && p->gate2_b && gate), "gated head needs gate params");
    size_t need = 0;
    const Saved s = carve_saved(saved, B, rna_dim, &need);
    if (saved_bytes < need) { set_error("head fwd: saved buffer %zu < %zu", saved_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    const Scratch w = carve_scratch(workspace, B, &need);
    if (workspace_bytes < need) { set_error("head fwd: workspace %zu < %zu", workspace_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    const int Kp = kpad(rna_dim);
    const uint32_t thresh = training ? drop_thresh(dropout_p) : 0;
    // training == B200SURV_HEAD_TRAIN_SEED_DEV: `seed` is the address of a uint64 in device memory
    const uint64_t *seed_dev = training == B200SURV_HEAD_TRAIN_SEED_DEV ? reinterpret_cast<const uint64_t *>(static_cast<uintptr_t>(seed)) : nullptr;
    const float inv_keep = 1.f / (1.f - dropout_p);
    const int bsl = bn_slices(B);
    int32_t rc;

    // bf16 operand copies (K padded to a multiple of 8 for the TMA row pitch): the batch (unless the caller staged it with
    // b200surv_head_stage_rna), then all weights in one launch
    if (!x_staged) cast_pad(rna, rna_dim, s.xb, Kp, B, rna_dim, Kp, st);
    {
        CastSegs cs;
        cs.n = 0;
        auto add = [&](const float *src, int lds, bf16 *dst, int ldd, int R, int C, int Cp) {
            cs.s[cs.n].src = src; cs.s[cs.n].dst = dst; cs.s[cs.n].lds = lds; cs.s[cs.n].ldd = ldd;
            cs.s[cs.n].R = R; cs.s[cs.n].C = C; cs.s[cs.n].Cp = Cp; ++cs.n;
        };
        add(p->rna0_w, rna_dim, s.w1b, Kp, H1, rna_dim, Kp);
        add(p->rna4_w, H1, s.w2b, H1, R1, H1, H1);
        if (gated) add(p->gate0_w, GZ, s.wg1b, GZP, GH, GZ, GZP);
        add(p->fus0_w, FEAT, s.wf1b, FEAT, H2, FEAT, FEAT);
        add(p->fus4_w, H2, s.wf2b, H2, F2N, H2, H2);
        cs.b0[0] = 0;
        for (int k = 0; k < cs.n; ++k)
            cs.b0[k + 1] = cs.b0[k] + ((cs.s[k].R + CAST_ROWS - 1) / CAST_ROWS) * ((cs.s[k].Cp + 255) / 256);
        // B200SURV_HEAD_SEED_ADVANCE: the device-resident dropout seed is bumped here, by the first kernel of the forward
        // pass -- every later kernel of this forward / backward pair reads the new value
        launch_chain(k_cast_multi, cs.b0[cs.n], 256, 0, st, cs, (seed_adv && seed_dev) ? const_cast<uint64_t *>(seed_dev) : nullptr);
    }

    // rna encoder: Linear(rna_dim, 512) -> BN -> ReLU -> Dropout -> Linear(512, 128) -> ReLU
    // Large batches: CTA pairs on 256 x 256 tiles, K split in two (64 pairs x 2 = 128 CTAs); the BatchNorm statistics kernel
    // adds the two slices and the bias.  Small batches: one 128-wide tile kernel writes h1 directly.
    const bool pair1 = B >= 512 && rna_dim >= 2048;
    if (pair1) {
        rc = gemm_bf16_ex(s.xb, Kp, 0, s.w1b, Kp, 0, (int)B, H1, rna_dim, w.slices, H1, nullptr, 0, nullptr, 0, w.slices, 512, 2, st);
        if (rc) return rc;
        launch_chain(k_bn_colstats<2>, dim3(H1 / BC, bsl), 256, 0, st, w.slices, w.slices + (size_t)B * H1, p->rna0_b, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, B, H1, 0, 0.f, 0, nullptr, 0, training ? 1 : 0, s.h1,
                                                             w.partial);
    } else {
        rc = gemm_bf16(s.xb, Kp, 0, s.w1b, Kp, 0, (int)B, H1, rna_dim, s.h1, H1, nullptr, 0, p->rna0_b, 0, nullptr, st);
        if (rc) return rc;
        if (training)
            launch_chain(k_bn_colstats<0>, dim3(H1 / BC, bsl), 256, 0, st, s.h1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, H1,
                                                                 0, 0.f, 0, nullptr, 0, 1, nullptr, w.partial);
    }
    launch_chain(k_bn_apply2, dim3(H1 / BC, (unsigned)((B + BR - 1) / BR)), 256, 0, st, s.h1, w.partial, bsl, p->bn1_w, p->bn1_b, B, H1, training,
                                                                              p->bn1_rm, p->bn1_rv, s.mu1, s.rstd1, thresh, inv_keep,
                                                                              seed, seed_dev, 1, s.a1, keep1);
    rc = gemm_bf16(s.a1, H1, 0, s.w2b, H1, 0, (int)B, R1, H1, s.r, R1, nullptr, 0, p->rna4_b, 1, nullptr, st);
    if (rc) return rc;

    // clinical encoder, masks, gate
    launch_chain(k_gate_prep, gs(B * GZP), 256, 0, st, ct_feat, s.r, clinical, mask, p->clin_w, p->clin_b, B, s.feat,
                                             gated ? s.z : nullptr, gated ? nullptr : s.fused);
    if (gated) {
        rc = gemm_bf16(s.z, GZP, 0, s.wg1b, GZP, 0, (int)B, GH, GZ, s.zh, GH, nullptr, 0, p->gate0_b, 1, nullptr, st);
        if (rc) return rc;
        launch_chain(k_gate_apply, gs(B * 32), 256, 0, st, s.zh, p->gate2_w, p->gate2_b, s.feat, B, s.gate, gate, s.fused);
    }
    // fusion: Linear(288, 256) -> BN -> ReLU -> Dropout -> Linear(256, 128) -> ReLU ; cox head
    rc = gemm_bf16(s.fused, FEAT, 0, s.wf1b, FEAT, 0, (int)B, H2, FEAT, s.h2, H2, nullptr, 0, p->fus0_b, 0, nullptr, st);
    if (rc) return rc;
    if (training)
        launch_chain(k_bn_colstats<0>, dim3(H2 / BC, bsl), 256, 0, st, s.h2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, H2, 0,
                                                             0.f, 0, nullptr, 0, 1, nullptr, w.partial);
    launch_chain(k_bn_apply2, dim3(H2 / BC, (unsigned)((B + BR - 1) / BR)), 256, 0, st, s.h2, w.partial, bsl, p->bn2_w, p->bn2_b, B, H2, training,
                                                                              p->bn2_rm, p->bn2_rv, s.mu2, s.rstd2, thresh, inv_keep,
                                                                              seed, seed_dev, 2, s.a2, keep2);
    rc = gemm_bf16(s.a2, H2, 0, s.wf2b, H2, 0, (int)B, F2N, H2, s.f2, F2N, nullptr, 0, p->fus4_b, 1, nullptr, st);
    if (rc) return rc;
    launch_chain(k_cox_head, gs(B * 32), 256, 0, st, s.f2, p->cox_w, p->cox_b, B, hazard);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_head_stage_rna(const float *rna, int64_t B, int32_t rna_dim, void *saved, size_t saved_bytes,
                                b200surv_stream_t stream) {
    B200_REQUIRE(rna && saved && B >= 1 && rna_dim >= 1, "arguments");
    size_t need = 0;
    const Saved s = carve_saved(saved, B, rna_dim, &need);
    if (saved_bytes < need) { set_error("head stage: saved buffer %zu < %zu", saved_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    cast_pad(rna, rna_dim, s.xb, kpad(rna_dim), B, rna_dim, kpad(rna_dim), as_stream(stream));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_head_stage_batch(const float *rna, int64_t B, int32_t rna_dim, void *saved, size_t saved_bytes,
                                  const float *const *copy_src, float *const *copy_dst, const int64_t *copy_elems,
                                  int32_t n_copies, b200surv_stream_t stream) {
    B200_REQUIRE(rna && saved && B >= 1 && rna_dim >= 1, "arguments");
    B200_REQUIRE(n_copies >= 0 && n_copies <= 3 && (n_copies == 0 || (copy_src && copy_dst && copy_elems)), "copies");
    size_t need = 0;
    const Saved s = carve_saved(saved, B, rna_dim, &need);
    if (saved_bytes < need) { set_error("head stage: saved buffer %zu < %zu", saved_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    StageCopies cp;
    cp.count = 0;
    for (int k = 0; k < n_copies; ++k) {
        if (copy_elems[k] <= 0) continue;
        B200_REQUIRE(copy_src[k] && copy_dst[k], "copy pointers");
        cp.src[cp.count] = copy_src[k]; cp.dst[cp.count] = copy_dst[k]; cp.n[cp.count] = copy_elems[k]; ++cp.count;
    }
    const int Kp = kpad(rna_dim);
    B200_REQUIRE(Kp % 4 == 0, "padded row length");
    const int gx = (Kp / 4 + 255) / 256;
    int64_t gy = (B + 1) / 2;                                // two rows per thread and round
    const int64_t cap = (int64_t)16 * num_sms() / gx + 1;
    if (gy > cap) gy = cap;
    const int ncopy = cp.count ? 2 * num_sms() : 0;
    launch_chain(k_stage_batch, (unsigned)(gx * gy + ncopy), 256, 0, as_stream(stream), rna, (int64_t)rna_dim, s.xb, (int64_t)Kp, B,
                 (int)rna_dim, Kp, gx, (int)gy, cp);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_head_bwd(const b200surv_head_params *p, const b200surv_head_grads *g, const float *d_hazard,
                          const float *d_gate, const float *clinical, const float *mask, int64_t B, int32_t rna_dim,
                          int32_t training, float dropout_p, uint64_t seed, float *d_ct_feat, const void *saved,
                          size_t saved_bytes, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(p && g && d_hazard && clinical && saved && workspace, "null pointer");
    B200_REQUIRE(B >= 1 && rna_dim >= 1, "B, rna_dim");
    training &= ~(B200SURV_HEAD_X_STAGED | B200SURV_HEAD_SEED_ADVANCE);
    const bool gated = mask != nullptr;
    size_t need = 0;
    const Saved s = carve_saved(const_cast<void *>(saved), B, rna_dim, &need);
    if (saved_bytes < need) { set_error("head bwd: saved buffer %zu < %zu", saved_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    const Scratch w = carve_scratch(workspace, B, &need);
    if (workspace_bytes < need) { set_error("head bwd: workspace %zu < %zu", workspace_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    const int Kp = kpad(rna_dim);
    const uint32_t thresh = training ? drop_thresh(dropout_p) : 0;
    // training == B200SURV_HEAD_TRAIN_SEED_DEV: `seed` is the address of a uint64 in device memory
    const uint64_t *seed_dev = training == B200SURV_HEAD_TRAIN_SEED_DEV ? reinterpret_cast<const uint64_t *>(static_cast<uintptr_t>(seed)) : nullptr;
    const float inv_keep = 1.f / (1.f - dropout_p);
    const int bsl = bn_slices(B);
    int32_t rc;

    // The caller's stream carries the chain of input-gradient GEMMs; the small weight-gradient GEMMs and the slice sums of
    // the fused reductions run on the side stream, forked / joined with events (works under stream capture as well).
    HeadLanes *lanes = head_lanes();
    cudaStream_t sd = lanes ? lanes->side : st;
    auto fork = [&]() -> int32_t {  // everything issued on `st` so far precedes what the side stream does next
        if (!lanes) return B200SURV_OK;
        B200_CHECK_CUDA(cudaEventRecord(lanes->fork, st));
        B200_CHECK_CUDA(cudaStreamWaitEvent(sd, lanes->fork, 0));
        return B200SURV_OK;
    };
    const int fsl = fslices_for(B);
    auto sums = [&](const double *partial, int ncols, SumSegs sg) {
        launch_chain(k_sum_slices, (ncols + 7) / 8, 256, 0, sd, partial, fsl, ncols, sg);
    };

    // ---- cox head: df2 (ReLU-masked, bf16) + partial sums of cox_head.weight / .bias and fusion.4.bias, one launch
    launch_chain(k_cox_bwd_fused, fsl, F2N, 0, st, d_hazard, s.f2, p->cox_w, B, w.df2, w.p_cox);
    if ((rc = fork())) return rc;
    {
        SumSegs sg; sg.n = 3;
        sg.out[0] = g->cox_w; sg.out[1] = g->fus4_b; sg.out[2] = g->cox_b; sg.out[3] = nullptr;
        sg.c0[0] = 0; sg.c0[1] = F2N; sg.c0[2] = 2 * F2N; sg.c0[3] = 2 * F2N + 1; sg.c0[4] = 2 * F2N + 1;
        sums(w.p_cox, 3 * F2N, sg);
    }
    // ---- fusion.4: dW = dF2^T a2 (side), dA2 = dF2 Wf2
    rc = gemm_wgrad(w.df2, F2N, s.a2, H2, F2N, H2, (int)B, g->fus4_w, H2, w.splitk, sd);
    if (rc) return rc;
    rc = gemm_bf16(w.df2, F2N, 0, s.wf2b, H2, 1, (int)B, H2, F2N, w.t0, H2, nullptr, 0, nullptr, 0, nullptr, st);  // t0 = dA2 [B][256]
    if (rc) return rc;
    // ---- fusion.1-3 (BN, ReLU, Dropout) backward -> dH2 (bf16); fusion.0.bias with it
    launch_chain(k_bn_colstats<1>, dim3(H2 / BC, bsl), 256, 0, st, w.t0, nullptr, nullptr, s.h2, s.mu2, s.rstd2, p->bn2_w, p->bn2_b, B, H2, thresh,
                                                         inv_keep, seed, seed_dev, 2, 1, nullptr, w.partial);
    launch_chain(k_bn_bwd_dx3, dim3(H2 / BC, (unsigned)((B + BR - 1) / BR)), 256, 0, st, w.t0, s.h2, w.partial, bsl, s.mu2, s.rstd2, p->bn2_w,
                                                                               p->bn2_b, B, H2, training, thresh, inv_keep, seed,
                                                                               seed_dev, 2, g->bn2_b, g->bn2_w, g->fus0_b, w.dh2);
    // ---- fusion.0: dW = dH2^T fused (side), dfused = dH2 Wf1
    if ((rc = fork())) return rc;
    rc = gemm_wgrad(w.dh2, H2, s.fused, FEAT, H2, FEAT, (int)B, g->fus0_w, FEAT, w.splitk, sd);
    if (rc) return rc;
    rc = gemm_bf16(w.dh2, H2, 0, s.wf1b, FEAT, 1, (int)B, FEAT, H2, w.t0, FEAT, nullptr, 0, nullptr, 0, nullptr, st);  // t0 = dfused
    if (rc) return rc;
    const float *dfeat = w.t0;
    const float *dz = nullptr;
    if (gated) {
        // ---- gate: softmax / scaling backward with the gate.2 gradients and gate.0.bias as partial sums; gate.0
        launch_chain(k_gate_bwd_fused, fsl, 256, 0, st, w.t0, s.feat, s.gate, s.zh, p->gate2_w, d_gate, B, w.t1, w.dzh, w.p_gate);
        if ((rc = fork())) return rc;
        {
            SumSegs sg; sg.n = 3;
            sg.out[0] = g->gate2_w; sg.out[1] = g->gate2_b; sg.out[2] = g->gate0_b; sg.out[3] = nullptr;
            sg.c0[0] = 0; sg.c0[1] = 3 * GH; sg.c0[2] = 3 * GH + 3; sg.c0[3] = GP_COLS; sg.c0[4] = GP_COLS;
            sums(w.p_gate, GP_COLS, sg);
        }
        rc = gemm_wgrad(w.dzh, GH, s.z, GZP, GH, GZ, (int)B, g->gate0_w, GZ, w.splitk, sd);
        if (rc) return rc;
        rc = gemm_bf16(w.dzh, GH, 0, s.wg1b, GZP, 1, (int)B, GZP, GH, w.t0, GZP, nullptr, 0, nullptr, 0, nullptr, st);  // t0 = dz
        if (rc) return rc;
        dfeat = w.t1;
        dz = w.t0;
    }
    // ---- masks, clinical encoder (its two gradients and rna_encoder.4.bias as partial sums)
    launch_chain(k_prep_bwd_fused, fsl, FEAT, 0, st, dfeat, dz, GZP, mask, s.r, clinical, p->clin_w, p->clin_b, B, d_ct_feat, w.dR, w.p_prep);
    if ((rc = fork())) return rc;
    {
        SumSegs sg; sg.n = 3;
        sg.out[0] = g->rna4_b; sg.out[1] = g->clin_w; sg.out[2] = g->clin_b; sg.out[3] = nullptr;
        sg.c0[0] = 0; sg.c0[1] = R1; sg.c0[2] = R1 + CL; sg.c0[3] = PP_COLS; sg.c0[4] = PP_COLS;
        sums(w.p_prep, PP_COLS, sg);
    }
    // ---- rna_encoder.4: dW = dR^T a1 (side), dA1 = dR W2
    rc = gemm_wgrad(w.dR, R1, s.a1, H1, R1, H1, (int)B, g->rna4_w, H1, w.splitk, sd);
    if (rc) return rc;
    rc = gemm_bf16(w.dR, R1, 0, s.w2b, H1, 1, (int)B, H1, R1, w.t0, H1, nullptr, 0, nullptr, 0, nullptr, st);  // t0 = dA1 [B][512]
    if (rc) return rc;
    // ---- rna_encoder.1-3 backward -> dH1 (bf16); rna_encoder.0.bias with it
    launch_chain(k_bn_colstats<1>, dim3(H1 / BC, bsl), 256, 0, st, w.t0, nullptr, nullptr, s.h1, s.mu1, s.rstd1, p->bn1_w, p->bn1_b, B, H1, thresh,
                                                         inv_keep, seed, seed_dev, 1, 1, nullptr, w.partial);
    launch_chain(k_bn_bwd_dx3, dim3(H1 / BC, (unsigned)((B + BR - 1) / BR)), 256, 0, st, w.t0, s.h1, w.partial, bsl, s.mu1, s.rstd1, p->bn1_w,
                                                                               p->bn1_b, B, H1, training, thresh, inv_keep, seed,
                                                                               seed_dev, 1, g->bn1_b, g->bn1_w, g->rna0_b, w.dh1);
    // ---- rna_encoder.0: dW1 [512][rna_dim] = dH1^T x  (the big one; x is an input, no dx): CTA pairs for large batches
    if (B >= 512 && rna_dim >= 2048)
        rc = gemm_bf16_ex(w.dh1, H1, 1, s.xb, Kp, 1, H1, rna_dim, (int)B, g->rna0_w, rna_dim, nullptr, 0, nullptr, 0, nullptr, 512, 0, st);
    else
        rc = gemm_wgrad(w.dh1, H1, s.xb, Kp, H1, rna_dim, (int)B, g->rna0_w, rna_dim, w.splitk, st);
    if (rc) return rc;
    if (lanes) {  // join: the caller's stream continues after the side stream's work
        B200_CHECK_CUDA(cudaEventRecord(lanes->join, sd));
        B200_CHECK_CUDA(cudaStreamWaitEvent(st, lanes->join, 0));
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_gate_entropy_fwd(const float *gate, int64_t B, float eps, float *out_loss, b200surv_stream_t stream) {
    B200_REQUIRE(gate && out_loss && B >= 1, "arguments");
    launch_chain(k_gate_entropy_fwd, 1, 1024, 0, as_stream(stream), gate, B, eps, out_loss);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}
int32_t b200surv_gate_entropy_bwd(const float *gate, const float *grad_out, int64_t B, float eps, float *d_gate,
                                  b200surv_stream_t stream) {
    B200_REQUIRE(gate && grad_out && d_gate && B >= 1, "arguments");
    launch_chain(k_gate_entropy_bwd, gs(3 * B), 256, 0, as_stream(stream), gate, grad_out, B, eps, d_gate);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // extern "C"
