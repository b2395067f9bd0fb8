// CT encoder feeding the fusion head: the reference's non-MONAI branch,
// scripts/training/partial_modality_training.py:179-190
//     Conv3d(1,32,3,s2,p1) BN3d ReLU  Conv3d(32,64,3,s2,p1) BN3d ReLU  Conv3d(64,128,3,s2,p1) BN3d ReLU  AdaptiveAvgPool3d(1)
// as B200 primitives behind the C ABI; the host layer (ctenc.py) strings them together.
//
// Layout: activations are CHANNELS-LAST, rows = (sample, z, y, x) of a layer's output grid, columns = channels, so that
//   * BatchNorm3d is a per-column normalisation of an [R][C] matrix (R = B * voxels),
//   * the 3x3x3 stride-2 convolutions with Cin >= 32 are GEMMs [R][27 Cin] x [27 Cin][Cout] on the tcgen05 kernel of
//     gemm_tc.cu: the patch matrix ("col", tap-major: column = tap * Cin + c, so every tap is one contiguous Cin-vector
//     of the source voxel) is written by k_im2col for a CHUNK of samples sized to stay in the 126 MB L2 between the
//     im2col store and the GEMM's TMA loads,
//   * the first convolution (Cin = 1, K = 27: too thin for the tensor pipe) is a direct CUDA-core kernel.
// Backward: BN/ReLU backward in two streaming passes (column sums, then dx as bf16), weight gradients = dx^T col on the
// GEMM (both operands MN-major), input gradients = dx W on the GEMM followed by a GATHER col2im (every input voxel pulls
// its <= 8 (tap, output voxel) contributions: no atomics, deterministic).  Every reduction sums in a fixed order.
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200surv {
// gemm_tc.cu
int32_t gemm_bf16(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int M, int N, int K,
                  float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias, int relu, float *splitk_ws,
                  cudaStream_t st);
int splitk_slices(int M, int N, int K, int *kb_per);
namespace {

using bf16 = __nv_bfloat16;
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;

struct Grid3 {            // input grid of one stride-2 convolution and its output grid
    int D, H, W, Do, Ho, Wo;
};
inline int out_dim(int d) { return (d - 1) / 2 + 1; }   // floor((d + 2*1 - 3) / 2) + 1
inline Grid3 make_grid(int D, int H, int W) { return Grid3{D, H, W, out_dim(D), out_dim(H), out_dim(W)}; }

inline unsigned blocks_for(int64_t items, int per_block, int max_per_sm) {
    int64_t g = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)max_per_sm * num_sms();
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ---------------------------------------------------------------- first convolution (Cin = 1), direct
// x fp32 [B][D][H][W]; w fp32 [Cout][27] (torch (Cout,1,3,3,3)); h fp32 [B*Do*Ho*Wo][Cout].  One thread = one output
// voxel x 8 channels (Cout/8 neighbouring threads share the voxel and write 32 contiguous bytes each).
// IDX: uint32_t when every index fits (64-bit divisions cost ~100 instructions each), else int64_t
template <typename IDX>
__global__ void __launch_bounds__(256)
k_conv_first_fwd(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, int64_t B,
                 Grid3 g, int Cout, float *__restrict__ h) {
    pdl_prologue();
    extern __shared__ float sw[];           // [27][Cout] then bias[Cout]
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) { const int c = i % Cout, t = i / Cout; sw[i] = w[c * 27 + t]; }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const IDX groups = (IDX)(Cout / 8);
    const IDX vox = (IDX)g.Do * g.Ho * g.Wo, total = (IDX)B * vox * groups;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int cg = (int)(i % groups);
        const IDX r = i / groups;
        const IDX b = r / vox;
        int v = (int)(r - b * vox);
        const int xo = v % g.Wo; v /= g.Wo;
        const int yo = v % g.Ho, zo = v / g.Ho;
        const float *xb = x + (int64_t)b * g.D * g.H * g.W;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = sw[27 * Cout + cg * 8 + j];
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
            const int z = 2 * zo - 1 + kz;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int y = 2 * yo - 1 + ky;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = 2 * xo - 1 + kx;
                    const bool in = z >= 0 && z < g.D && y >= 0 && y < g.H && xx >= 0 && xx < g.W;
                    const float xv = in ? __ldg(xb + ((int64_t)z * g.H + y) * g.W + xx) : 0.f;
                    const float4 w0 = *reinterpret_cast<const float4 *>(sw + (kz * 9 + ky * 3 + kx) * Cout + cg * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(sw + (kz * 9 + ky * 3 + kx) * Cout + cg * 8 + 4);
                    acc[0] += xv * w0.x; acc[1] += xv * w0.y; acc[2] += xv * w0.z; acc[3] += xv * w0.w;
                    acc[4] += xv * w1.x; acc[5] += xv * w1.y; acc[6] += xv * w1.z; acc[7] += xv * w1.w;
                }
            }
        }
        float4 *o = reinterpret_cast<float4 *>(h + (int64_t)r * Cout + cg * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// weight gradient of the first convolution: dw[c][tap] = sum_r dx[r][c] * patch_r[tap].  Thread = (row lane, channel):
// a warp owns a contiguous run of output rows (voxel coordinates advance incrementally: no division in the loop) and
// its lanes the channels, so the 27 patch values are warp-uniform broadcast loads; 27 accumulators per thread.
// partial[cta][Cout][27] (fp32) is summed in CTA order by k_sum_slices.
constexpr int WG_THREADS = 1024;
__global__ void __launch_bounds__(WG_THREADS, 1)
k_conv_first_wgrad(const float *__restrict__ x, const bf16 *__restrict__ dx, int64_t B, Grid3 g, int Cout,
                   float *__restrict__ partial) {
    pdl_prologue();
    __shared__ float red[WG_THREADS * 9];
    const int c = threadIdx.x % Cout, lane_r = threadIdx.x / Cout, lanes = WG_THREADS / Cout;
    const int64_t vox = (int64_t)g.Do * g.Ho * g.Wo, R = B * vox;
    const int64_t per = (R + (int64_t)gridDim.x * lanes - 1) / ((int64_t)gridDim.x * lanes);
    const int64_t r0 = ((int64_t)blockIdx.x * lanes + lane_r) * per, r1 = min(R, r0 + per);
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
    if (r0 < r1) {
        int64_t b = r0 / vox;
        int v = (int)(r0 - b * vox);
        int xo = v % g.Wo; v /= g.Wo;
        int yo = v % g.Ho, zo = v / g.Ho;
        const int64_t plane = (int64_t)g.H * g.W;
        const float *xb = x + b * g.D * plane;
        for (int64_t r = r0; r < r1; ++r) {
            const float d = __bfloat162float(dx[r * Cout + c]);
            const float *x0 = xb + (int64_t)(2 * zo - 1) * plane + (int64_t)(2 * yo - 1) * g.W + (2 * xo - 1);
#pragma unroll
            for (int kz = 0; kz < 3; ++kz) {
                const bool inz = (unsigned)(2 * zo - 1 + kz) < (unsigned)g.D;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const bool iny = inz && (unsigned)(2 * yo - 1 + ky) < (unsigned)g.H;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const bool in = iny && (unsigned)(2 * xo - 1 + kx) < (unsigned)g.W;
                        const float xv = in ? __ldg(x0 + kz * plane + ky * g.W + kx) : 0.f;
                        acc[kz * 9 + ky * 3 + kx] += d * xv;
                    }
                }
            }
            if (++xo == g.Wo) {
                xo = 0;
                if (++yo == g.Ho) {
                    yo = 0;
                    if (++zo == g.Do) { zo = 0; xb += g.D * plane; }
                }
            }
        }
    }
#pragma unroll
    for (int round = 0; round < 3; ++round) {       // nine taps at a time through 36 KB of shared memory
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 9; ++t) red[threadIdx.x * 9 + t] = acc[round * 9 + t];
        __syncthreads();
        for (int i = threadIdx.x; i < Cout * 9; i += WG_THREADS) {
            const int cc = i / 9, t = i % 9;
            float sum = 0.f;
            for (int l = 0; l < lanes; ++l) sum += red[(l * Cout + cc) * 9 + t];
            partial[(size_t)blockIdx.x * Cout * 27 + cc * 27 + round * 9 + t] = sum;
        }
    }
}
// The reference's width (Cout = 32): lanes = the 27 taps, a warp = a contiguous run of output rows.  Per row one
// (predicated) load of the lane's own tap and FOUR warp-uniform 16-byte loads of the row's 32 bf16 dx values, then 32
// FMAs into per-channel accumulators: 5 loads per row and warp instead of 28.
__global__ void __launch_bounds__(WG_THREADS, 1)
k_conv_first_wgrad32(const float *__restrict__ x, const bf16 *__restrict__ dx, int64_t B, Grid3 g,
                     float *__restrict__ partial) {
    pdl_prologue();
    __shared__ float red[WG_THREADS * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = WG_THREADS / 32;
    const int tap = lane < 27 ? lane : 26;
    const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
    const int64_t vox = (int64_t)g.Do * g.Ho * g.Wo, R = B * vox;
    const int64_t per = (R + (int64_t)gridDim.x * warps - 1) / ((int64_t)gridDim.x * warps);
    const int64_t r0 = ((int64_t)blockIdx.x * warps + warp) * per, r1 = min(R, r0 + per);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    if (r0 < r1) {
        int64_t b = r0 / vox;
        int v = (int)(r0 - b * vox);
        int xo = v % g.Wo; v /= g.Wo;
        int yo = v % g.Ho, zo = v / g.Ho;
        const int64_t plane = (int64_t)g.H * g.W;
        const float *xb = x + b * g.D * plane;
        for (int64_t r = r0; r < r1; ++r) {
            const int z = 2 * zo - 1 + kz, y = 2 * yo - 1 + ky, xx = 2 * xo - 1 + kx;
            const bool in = (unsigned)z < (unsigned)g.D && (unsigned)y < (unsigned)g.H && (unsigned)xx < (unsigned)g.W;
            const float xv = in ? __ldg(xb + z * plane + (int64_t)y * g.W + xx) : 0.f;
            const uint4 *dr = reinterpret_cast<const uint4 *>(dx + r * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 u = __ldg(dr + q);
                const unsigned w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[q * 8 + 2 * j] += xv * __uint_as_float(w4[j] << 16);
                    acc[q * 8 + 2 * j + 1] += xv * __uint_as_float(w4[j] & 0xffff0000u);
                }
            }
            if (++xo == g.Wo) {
                xo = 0;
                if (++yo == g.Ho) {
                    yo = 0;
                    if (++zo == g.Do) { zo = 0; xb += g.D * plane; }
                }
            }
        }
    }
#pragma unroll
    for (int round = 0; round < 4; ++round) {       // eight channels at a time through 32 KB of shared memory
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 8; ++c) red[threadIdx.x * 8 + c] = acc[round * 8 + c];
        __syncthreads();
        for (int i = threadIdx.x; i < 27 * 8; i += WG_THREADS) {
            const int t = i / 8, c = i % 8;
            float sum = 0.f;
            for (int w = 0; w < warps; ++w) sum += red[(w * 32 + t) * 8 + c];
            partial[(size_t)blockIdx.x * 32 * 27 + (round * 8 + c) * 27 + t] = sum;
        }
    }
}
// out[i] = sum over the slices, in a fixed order: 32 elements x 8 slice lanes per CTA
__global__ void __launch_bounds__(256)
k_sum_slices(const float *__restrict__ part, int slices, int64_t elems, float *__restrict__ out) {
    pdl_prologue();
    __shared__ double sh[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + tx;
    double sum = 0.0;
    if (i < elems)
        for (int k = ty; k < slices; k += 8) sum += part[(size_t)k * elems + i];
    sh[ty][tx] = sum;
    __syncthreads();
    if (ty == 0 && i < elems) {
#pragma unroll
        for (int k = 1; k < 8; ++k) sum += sh[k][tx];
        out[i] = (float)sum;
    }
}

// ---------------------------------------------------------------- im2col / col2im (stride 2, pad 1, 3x3x3)
// a bf16 [Bc*D*H*W][C] -> col bf16 [Bc*Do*Ho*Wo][27*C], column = tap*C + c.  A WARP copies one output row's 27*C
// values as 16-byte vectors (lane j <-> vector j: fully coalesced stores; a tap's C channels are one contiguous run of
// the source voxel), so the row's voxel coordinates are decomposed once per 27*C/8 vectors; CV = C/8 (0: run time).
template <int CV>
__global__ void __launch_bounds__(256)
k_im2col(const bf16 *__restrict__ a, int64_t Bc, Grid3 g, int C, bf16 *__restrict__ col) {
    pdl_prologue();
    const int cv = CV ? CV : C / 8, nvec = 27 * cv;
    const int lane = threadIdx.x & 31;
    const int64_t vox = (int64_t)g.Do * g.Ho * g.Wo, rows = Bc * vox, vin = (int64_t)g.D * g.H * g.W;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
        const int64_t b = r / vox;
        int v = (int)(r - b * vox);
        const int xo = v % g.Wo; v /= g.Wo;
        const int yo = v % g.Ho, zo = v / g.Ho;
        const uint4 *src = reinterpret_cast<const uint4 *>(a + b * vin * C);
        uint4 *dst = reinterpret_cast<uint4 *>(col + r * 27 * C);
        for (int j = lane; j < nvec; j += 32) {
            const int tap = j / cv, c8 = j - tap * cv;
            const int kz = tap / 9, ky = (tap - kz * 9) / 3, kx = tap - kz * 9 - ky * 3;
            const int z = 2 * zo - 1 + kz, y = 2 * yo - 1 + ky, xx = 2 * xo - 1 + kx;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if ((unsigned)z < (unsigned)g.D && (unsigned)y < (unsigned)g.H && (unsigned)xx < (unsigned)g.W)
                val = __ldg(src + (((int64_t)z * g.H + y) * g.W + xx) * cv + c8);
            dst[j] = val;
        }
    }
}
__device__ __forceinline__ void add_bf16x8(float *acc, uint4 v) {
    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(p[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
}
// dcol bf16 [Bc*Do*Ho*Wo][27*C] -> da fp32 [Bc*D*H*W][C]: input voxel (z,y,x) receives tap (kz,ky,kx) of output voxel
// ((z+1-kz)/2, ...) whenever that is an integer inside the output grid.
template <typename IDX>
__global__ void __launch_bounds__(256)
k_col2im(const bf16 *__restrict__ dcol, int64_t Bc, Grid3 g, int C, float *__restrict__ da) {
    pdl_prologue();
    const IDX cv = (IDX)(C / 8);
    const IDX vin = (IDX)g.D * g.H * g.W, total = (IDX)Bc * vin * cv;
    const int64_t vox = (int64_t)g.Do * g.Ho * g.Wo;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        const IDX p = i / cv;
        const IDX b = p / vin;
        int v = (int)(p - b * vin);
        const int xx = v % g.W; v /= g.W;
        const int y = v % g.H, z = v / g.H;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int kz = (z + 1) & 1; kz < 3; kz += 2) {
            const int zo = (z + 1 - kz) / 2;
            if (z + 1 - kz < 0 || zo >= g.Do) continue;
            for (int ky = (y + 1) & 1; ky < 3; ky += 2) {
                const int yo = (y + 1 - ky) / 2;
                if (y + 1 - ky < 0 || yo >= g.Ho) continue;
                for (int kx = (xx + 1) & 1; kx < 3; kx += 2) {
                    const int xo = (xx + 1 - kx) / 2;
                    if (xx + 1 - kx < 0 || xo >= g.Wo) continue;
                    const int64_t r = b * vox + ((int64_t)zo * g.Ho + yo) * g.Wo + xo;
                    const int tap = kz * 9 + ky * 3 + kx;
                    add_bf16x8(acc, __ldg(reinterpret_cast<const uint4 *>(dcol + (r * 27 + tap) * C) + c8));
                }
            }
        }
        float4 *o = reinterpret_cast<float4 *>(da + (int64_t)p * C + c8 * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ---------------------------------------------------------------- weights: torch (Cout,Cin,3,3,3) <-> tap-major
__global__ void k_weight_pack(const float *__restrict__ w, int Cout, int Cin, bf16 *__restrict__ wr) {
    pdl_prologue();
    const int total = Cout * Cin * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % Cin, tap = (i / Cin) % 27, o = i / (27 * Cin);       // i indexes wr
        wr[i] = __float2bfloat16_rn(w[((size_t)o * Cin + c) * 27 + tap]);
    }
}
__global__ void k_weight_unpack(const float *__restrict__ dwr, int slices, int Cout, int Cin, int accumulate,
                                float *__restrict__ dw) {
    pdl_prologue();
    const int total = Cout * Cin * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % 27, c = (i / 27) % Cin, o = i / (27 * Cin);        // i indexes dw
        double s = accumulate ? (double)dw[i] : 0.0;
        for (int k = 0; k < slices; ++k) s += dwr[(size_t)k * total + ((size_t)o * 27 + tap) * Cin + c];
        dw[i] = (float)s;
    }
}

// ---------------------------------------------------------------- BatchNorm3d on [R][C] (channels-last), C | 1024
// Column sums of an [R][C] fp32 matrix read as float4: thread t always sees channels (4t) % C .. +3.
// MODE 0: (sum x, sum x^2).  MODE 1: dy = [bn(x) > 0] * dA -> (sum dy, sum dy * xhat).
template <int MODE>
__global__ void __launch_bounds__(256)
k_bn_colsums(const float *__restrict__ x, const float *__restrict__ dA, const float *__restrict__ mu,
             const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ beta, int64_t R,
             int C, double *__restrict__ partial) {
    pdl_prologue();
    __shared__ double sh[256][8];
    const int t = threadIdx.x, c0 = (4 * t) % C;
    const int64_t nvec = R * C / 4;
    const int64_t per = ((nvec + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;     // keeps t <-> channel fixed
    const int64_t v0 = (int64_t)blockIdx.x * per, v1 = min(nvec, v0 + per);
    float m[4] = {0, 0, 0, 0}, rs[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0}, be[4] = {0, 0, 0, 0};
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { m[j] = mu[c0 + j]; rs[j] = rstd[c0 + j]; ga[j] = gamma[c0 + j]; be[j] = beta[c0 + j]; }
    }
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    int64_t v = v0 + t;
    while (v < v1) {
        float f0[4] = {0, 0, 0, 0}, f1[4] = {0, 0, 0, 0};
        for (int it = 0; it < 16 && v < v1; ++it, v += 256) {      // short fp32 runs, folded into fp64
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + v);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { f0[j] += xa[j]; f1[j] += xa[j] * xa[j]; }
            } else {
                const float4 dv = __ldg(reinterpret_cast<const float4 *>(dA) + v);
                const float da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xa[j] - m[j]) * rs[j];
                    const float dy = (xh * ga[j] + be[j] > 0.f) ? da[j] : 0.f;
                    f0[j] += dy; f1[j] += dy * xh;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0[j] += f0[j]; s1[j] += f1[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh[t][j] = s0[j]; sh[t][4 + j] = s1[j]; }
    __syncthreads();
    if (t < C) {
        const int q = C / 4;                  // threads t' = t/4 + k*q hold channel t
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < 256 / q; ++k) { a0 += sh[t / 4 + k * q][t % 4]; a1 += sh[t / 4 + k * q][4 + t % 4]; }
        partial[((size_t)blockIdx.x * 2 + 0) * C + t] = a0;
        partial[((size_t)blockIdx.x * 2 + 1) * C + t] = a1;
    }
}
// one WARP per channel, lanes over the partials in a fixed order (deterministic); grid = C / 8 blocks of 256 threads
__device__ __forceinline__ void part_sums(const double *__restrict__ partial, int nparts, int C, int c, int lane,
                                          double &s0, double &s1) {
    s0 = 0.0; s1 = 0.0;
    for (int k = lane; k < nparts; k += 32) { s0 += partial[((size_t)k * 2 + 0) * C + c]; s1 += partial[((size_t)k * 2 + 1) * C + c]; }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
}
__global__ void __launch_bounds__(256)
k_bn_finalize(const double *__restrict__ partial, int nparts, int64_t R, int C, int training,
              float *__restrict__ run_mean, float *__restrict__ run_var, float *__restrict__ mu,
              float *__restrict__ rstd) {
    pdl_prologue();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    if (!training) {
        if (lane == 0) { mu[c] = run_mean[c]; rstd[c] = rsqrtf(run_var[c] + BN_EPS); }
        return;
    }
    double s, ss;
    part_sums(partial, nparts, C, c, lane, s, ss);
    if (lane != 0) return;
    const double mean = s / (double)R;
    double var = ss / (double)R - mean * mean;
    if (var < 0.0) var = 0.0;
    mu[c] = (float)mean;
    rstd[c] = (float)(1.0 / sqrt(var + (double)BN_EPS));
    if (run_mean) {
        const double unb = R > 1 ? var * ((double)R / (double)(R - 1)) : var;
        run_mean[c] = (1.f - BN_MOM) * run_mean[c] + BN_MOM * (float)mean;
        run_var[c] = (1.f - BN_MOM) * run_var[c] + BN_MOM * (float)unb;
    }
}
// sums of the backward pass: sdy, sdyx (kept for k_bn_dx), dgamma = sdyx, dbeta = sdy, conv-bias gradient
// = sum_r dx = (train ? 0 : gamma rstd sdy)
__global__ void __launch_bounds__(256)
k_bn_bwd_finalize(const double *__restrict__ partial, int nparts, int C, int training, const float *__restrict__ gamma,
                  const float *__restrict__ rstd, float *__restrict__ sums, float *__restrict__ dgamma,
                  float *__restrict__ dbeta, float *__restrict__ dbias) {
    pdl_prologue();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s, sx;
    part_sums(partial, nparts, C, c, lane, s, sx);
    if (lane != 0) return;
    sums[c] = (float)s; sums[C + c] = (float)sx;
    if (dgamma) dgamma[c] = (float)sx;
    if (dbeta) dbeta[c] = (float)s;
    if (dbias) dbias[c] = training ? 0.f : gamma[c] * rstd[c] * (float)s;
}
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<unsigned *>(&lo); r.y = *reinterpret_cast<unsigned *>(&hi);
    return r;
}
// y bf16 = relu(bn(x))
__global__ void __launch_bounds__(256)
k_bn_relu(const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ rstd,
          const float *__restrict__ gamma, const float *__restrict__ beta, int64_t R, int C, bf16 *__restrict__ y) {
    pdl_prologue();
    const int64_t nvec = R * C / 4;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)((4 * v) & (C - 1));      // C is a power of two
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + v);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = fmaxf((xa[j] - mu[c0 + j]) * rstd[c0 + j] * gamma[c0 + j] + beta[c0 + j], 0.f);
        reinterpret_cast<uint2 *>(y)[v] = pack_bf16x4(o[0], o[1], o[2], o[3]);
    }
}
// feat[b][c] = mean_v relu(bn(x[b*V + v][c]))   (BN + ReLU + AdaptiveAvgPool3d(1) of the last stage).  One CTA per
// sample: thread = (voxel lane, channel), 1024 / C voxel lanes, summed in a fixed order.
__global__ void __launch_bounds__(1024)
k_bn_relu_pool(const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ rstd,
               const float *__restrict__ gamma, const float *__restrict__ beta, int64_t B, int V, int C,
               float *__restrict__ feat) {
    pdl_prologue();
    __shared__ float sh[1024];
    const int c = threadIdx.x % C, lane_v = threadIdx.x / C, lanes = 1024 / C;
    const float m = mu[c], sc = rstd[c] * gamma[c], be = beta[c];
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        float s = 0.f;
        for (int v = lane_v; v < V; v += lanes) s += fmaxf((x[(b * V + v) * C + c] - m) * sc + be, 0.f);
        __syncthreads();
        sh[threadIdx.x] = s;
        __syncthreads();
        if (lane_v == 0) {
            for (int l = 1; l < lanes; ++l) s += sh[l * C + c];
            feat[b * C + c] = s / (float)V;
        }
    }
}
__global__ void k_pool_bwd(const float *__restrict__ dfeat, int64_t B, int V, int C, float *__restrict__ dA) {
    pdl_prologue();
    const int64_t total = B * V * C;
    const float inv = 1.f / (float)V;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t b = i / ((int64_t)V * C);
        dA[i] = dfeat[b * C + c] * inv;
    }
}
// dx bf16 = gamma rstd (dy - sdy/R - xhat sdyx/R) in training, gamma rstd dy in eval; dy = [bn(x) > 0] dA
__global__ void __launch_bounds__(256)
k_bn_dx(const float *__restrict__ x, const float *__restrict__ dA, const float *__restrict__ mu,
        const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ beta,
        const float *__restrict__ sums, int64_t R, int C, int training, bf16 *__restrict__ dx) {
    pdl_prologue();
    const int64_t nvec = R * C / 4;
    const float invR = 1.f / (float)R;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)((4 * v) & (C - 1));      // C is a power of two
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + v);
        const float4 dv = __ldg(reinterpret_cast<const float4 *>(dA) + v);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            const float xh = (xa[j] - mu[c]) * rstd[c];
            float d = (xh * gamma[c] + beta[c] > 0.f) ? da[j] : 0.f;
            if (training) d = d - sums[c] * invR - xh * sums[C + c] * invR;
            o[j] = d * gamma[c] * rstd[c];
        }
        reinterpret_cast<uint2 *>(dx)[v] = pack_bf16x4(o[0], o[1], o[2], o[3]);
    }
}

inline bool pow2_channels(int C) { return C >= 8 && C <= 256 && (C & (C - 1)) == 0; }
constexpr int BN_PARTS_PER_SM = 4;
inline int bn_parts(int64_t R, int C) {
    int64_t p = (R * C / 4 + 1023) / 1024;      // >= 4 float4 per thread
    const int64_t cap = (int64_t)BN_PARTS_PER_SM * num_sms();
    if (p > cap) p = cap;
    if (p < 1) p = 1;
    return (int)p;
}

}  // namespace
}  // namespace b200surv

using namespace b200surv;

extern "C" {

int32_t b200surv_ct_conv_first_fwd(const float *x, const float *w, const float *bias, int64_t B, int32_t D, int32_t H,
                                   int32_t W, int32_t Cout, float *h, b200surv_stream_t stream) {
    B200_REQUIRE(x && w && h, "null pointer");
    B200_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "shape");
    B200_REQUIRE(Cout >= 8 && Cout % 8 == 0 && Cout <= 256, "Cout must be a multiple of 8, <= 256");
    const Grid3 g = make_grid(D, H, W);
    const int64_t total = B * (int64_t)g.Do * g.Ho * g.Wo * (Cout / 8);
    if (total < ((int64_t)1 << 30))
        launch_chain(k_conv_first_fwd<uint32_t>, blocks_for(total, 256, 16), 256, (size_t)28 * Cout * sizeof(float), as_stream(stream), x, w, bias, B, g, Cout, h);
    else
        launch_chain(k_conv_first_fwd<int64_t>, blocks_for(total, 256, 16), 256, (size_t)28 * Cout * sizeof(float), as_stream(stream), x, w, bias, B, g, Cout, h);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

size_t b200surv_ct_workspace_bytes(void) {
    // the larger of: BN column-sum partials [parts][2][256] fp64; first-conv weight-gradient partials [ctas][256*27] fp32
    const size_t bn = (size_t)BN_PARTS_PER_SM * num_sms() * 2 * 256 * sizeof(double) + 2 * 256 * sizeof(float);
    const size_t wg = (size_t)2 * num_sms() * 256 * 27 * sizeof(float);
    return align_up(bn > wg ? bn : wg, 256);
}

int32_t b200surv_ct_conv_first_wgrad(const float *x, const void *dx_bf16, int64_t B, int32_t D, int32_t H, int32_t W,
                                     int32_t Cout, float *dw, void *workspace, size_t workspace_bytes,
                                     b200surv_stream_t stream) {
    B200_REQUIRE(x && dx_bf16 && dw && workspace, "null pointer");
    B200_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "shape");
    B200_REQUIRE(Cout >= 8 && Cout <= 256 && WG_THREADS % Cout == 0, "Cout must be a power of two in [8, 256]");
    if (workspace_bytes < b200surv_ct_workspace_bytes()) { set_error("ct: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    const Grid3 g = make_grid(D, H, W);
    const int64_t R = B * (int64_t)g.Do * g.Ho * g.Wo;
    const int lanes = WG_THREADS / Cout;
    int ctas = num_sms();       // 64 registers x 1024 threads: one CTA per SM
    if ((int64_t)ctas * lanes > R) ctas = (int)((R + lanes - 1) / lanes);
    float *partial = static_cast<float *>(workspace);
    cudaStream_t st = as_stream(stream);
    if (Cout == 32) {
        ctas = num_sms();
        if ((int64_t)ctas * (WG_THREADS / 32) > R) ctas = (int)((R + WG_THREADS / 32 - 1) / (WG_THREADS / 32));
        launch_chain(k_conv_first_wgrad32, ctas, WG_THREADS, 0, st, x, static_cast<const bf16 *>(dx_bf16), B, g, partial);
    } else {
        launch_chain(k_conv_first_wgrad, ctas, WG_THREADS, 0, st, x, static_cast<const bf16 *>(dx_bf16), B, g, Cout, partial);
    }
    launch_chain(k_sum_slices, (Cout * 27 + 31) / 32, 256, 0, st, partial, ctas, (int64_t)Cout * 27, dw);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_im2col(const void *a_bf16, int64_t Bc, int32_t D, int32_t H, int32_t W, int32_t C, void *col_bf16,
                           b200surv_stream_t stream) {
    B200_REQUIRE(a_bf16 && col_bf16, "null pointer");
    B200_REQUIRE(Bc >= 1 && D >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "shape (C multiple of 8)");
    const Grid3 g = make_grid(D, H, W);
    const int64_t rows = Bc * (int64_t)g.Do * g.Ho * g.Wo;       // one warp per row, 8 warps per CTA
    const unsigned grid = blocks_for(rows, 8 * 4, 16);
    const bf16 *ap = static_cast<const bf16 *>(a_bf16);
    bf16 *cp = static_cast<bf16 *>(col_bf16);
    if (C == 32) launch_chain(k_im2col<4>, grid, 256, 0, as_stream(stream), ap, Bc, g, C, cp);
    else if (C == 64) launch_chain(k_im2col<8>, grid, 256, 0, as_stream(stream), ap, Bc, g, C, cp);
    else launch_chain(k_im2col<0>, grid, 256, 0, as_stream(stream), ap, Bc, g, C, cp);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_col2im(const void *dcol_bf16, int64_t Bc, int32_t D, int32_t H, int32_t W, int32_t C, float *da,
                           b200surv_stream_t stream) {
    B200_REQUIRE(dcol_bf16 && da, "null pointer");
    B200_REQUIRE(Bc >= 1 && D >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "shape (C multiple of 8)");
    const Grid3 g = make_grid(D, H, W);
    const int64_t total = Bc * (int64_t)D * H * W * (C / 8);
    if (total < ((int64_t)1 << 30))
        launch_chain(k_col2im<uint32_t>, blocks_for(total, 256 * 2, 16), 256, 0, as_stream(stream), static_cast<const bf16 *>(dcol_bf16), Bc, g, C, da);
    else
        launch_chain(k_col2im<int64_t>, blocks_for(total, 256 * 2, 16), 256, 0, as_stream(stream), static_cast<const bf16 *>(dcol_bf16), Bc, g, C, da);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_weight_pack(const float *w, int32_t Cout, int32_t Cin, void *wr_bf16, b200surv_stream_t stream) {
    B200_REQUIRE(w && wr_bf16 && Cout >= 1 && Cin >= 1, "arguments");
    launch_chain(k_weight_pack, (Cout * Cin * 27 + 255) / 256, 256, 0, as_stream(stream), w, Cout, Cin, static_cast<bf16 *>(wr_bf16));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_weight_unpack(const float *dwr_slices, int32_t slices, int32_t Cout, int32_t Cin, int32_t accumulate,
                                  float *dw, b200surv_stream_t stream) {
    B200_REQUIRE(dwr_slices && dw && slices >= 1 && Cout >= 1 && Cin >= 1, "arguments");
    launch_chain(k_weight_unpack, (Cout * Cin * 27 + 255) / 256, 256, 0, as_stream(stream), dwr_slices, slices, Cout, Cin, accumulate, dw);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_bn_stats(const float *x, int64_t R, int32_t C, int32_t training, float *run_mean, float *run_var,
                             float *mu, float *rstd, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(x && mu && rstd && workspace, "null pointer");
    B200_REQUIRE(R >= 1 && pow2_channels(C), "C must be a power of two in [8, 256]");
    B200_REQUIRE(training || (run_mean && run_var), "eval mode needs the running statistics");
    if (workspace_bytes < b200surv_ct_workspace_bytes()) { set_error("ct: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    double *partial = static_cast<double *>(workspace);
    const int parts = bn_parts(R, C);
    if (training)
        launch_chain(k_bn_colsums<0>, parts, 256, 0, st, x, nullptr, nullptr, nullptr, nullptr, nullptr, R, C, partial);
    launch_chain(k_bn_finalize, (C + 7) / 8, 256, 0, st, partial, parts, R, C, training, run_mean, run_var, mu, rstd);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_bn_relu(const float *x, const float *mu, const float *rstd, const float *gamma, const float *beta,
                            int64_t R, int32_t C, void *y_bf16, b200surv_stream_t stream) {
    B200_REQUIRE(x && mu && rstd && gamma && beta && y_bf16, "null pointer");
    B200_REQUIRE(R >= 1 && pow2_channels(C), "C must be a power of two in [8, 256]");
    launch_chain(k_bn_relu, blocks_for(R * C / 4, 256 * 4, 16), 256, 0, as_stream(stream), x, mu, rstd, gamma, beta, R, C,
                                                                             static_cast<bf16 *>(y_bf16));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_bn_relu_pool(const float *x, const float *mu, const float *rstd, const float *gamma,
                                 const float *beta, int64_t B, int32_t V, int32_t C, float *feat,
                                 b200surv_stream_t stream) {
    B200_REQUIRE(x && mu && rstd && gamma && beta && feat, "null pointer");
    B200_REQUIRE(B >= 1 && V >= 1 && pow2_channels(C), "shape (C a power of two in [8, 256])");
    launch_chain(k_bn_relu_pool, blocks_for(B, 1, 2), 1024, 0, as_stream(stream), x, mu, rstd, gamma, beta, B, V, C, feat);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_pool_bwd(const float *dfeat, int64_t B, int32_t V, int32_t C, float *dA, b200surv_stream_t stream) {
    B200_REQUIRE(dfeat && dA && B >= 1 && V >= 1 && C >= 1, "arguments");
    launch_chain(k_pool_bwd, blocks_for(B * V * C, 256 * 4, 16), 256, 0, as_stream(stream), dfeat, B, V, C, dA);
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

int32_t b200surv_ct_bn_bwd(const float *x, const float *dA, const float *mu, const float *rstd, const float *gamma,
                           const float *beta, int64_t R, int32_t C, int32_t training, void *dx_bf16, float *dgamma,
                           float *dbeta, float *dbias, void *workspace, size_t workspace_bytes,
                           b200surv_stream_t stream) {
    B200_REQUIRE(x && dA && mu && rstd && gamma && beta && dx_bf16 && workspace, "null pointer");
    B200_REQUIRE(R >= 1 && pow2_channels(C), "C must be a power of two in [8, 256]");
    if (workspace_bytes < b200surv_ct_workspace_bytes()) { set_error("ct: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    double *partial = static_cast<double *>(workspace);
    const int parts = bn_parts(R, C);
    // the 2 C fp32 column sums live behind the partials (b200surv_ct_workspace_bytes() reserves them)
    float *sums = reinterpret_cast<float *>(partial + (size_t)parts * 2 * C);
    B200_REQUIRE((size_t)parts * 2 * C * sizeof(double) + 2 * C * sizeof(float) <= workspace_bytes, "workspace layout");
    launch_chain(k_bn_colsums<1>, parts, 256, 0, st, x, dA, mu, rstd, gamma, beta, R, C, partial);
    launch_chain(k_bn_bwd_finalize, (C + 7) / 8, 256, 0, st, partial, parts, C, training, gamma, rstd, sums, dgamma, dbeta, dbias);
    launch_chain(k_bn_dx, blocks_for(R * C / 4, 256 * 4, 16), 256, 0, st, x, dA, mu, rstd, gamma, beta, sums, R, C, training,
                                                               static_cast<bf16 *>(dx_bf16));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}


// ---------------------------------------------------------------- the whole encoder in one call each way
// (the reference's fixed architecture 1 -> 32 -> 64 -> 128; ~35 launches forward, ~45 backward, no host work between)
}  // extern "C"
namespace {
constexpr size_t COL_BYTES_IN_L2 = (size_t)64 << 20;   // patch-matrix bytes per chunk of samples: half of the 126 MB L2
struct CtGeom {
    int64_t B;
    Grid3 g[3];
    int Cin[3], Cout[3], K[3];
    int64_t vin[3], vox[3], R[3], chunk[3];
};
CtGeom ct_geom(int64_t B, int D, int H, int W) {
    CtGeom q;
    q.B = B;
    const int ch[4] = {1, 32, 64, 128};
    int d[3] = {D, H, W};
    for (int s = 0; s < 3; ++s) {
        q.g[s] = make_grid(d[0], d[1], d[2]);
        q.Cin[s] = ch[s]; q.Cout[s] = ch[s + 1]; q.K[s] = 27 * ch[s];
        q.vin[s] = (int64_t)d[0] * d[1] * d[2];
        q.vox[s] = (int64_t)q.g[s].Do * q.g[s].Ho * q.g[s].Wo;
        q.R[s] = B * q.vox[s];
        int64_t c = (int64_t)(COL_BYTES_IN_L2 / ((size_t)q.vox[s] * q.K[s] * 2));
        q.chunk[s] = c < 1 ? 1 : (c > B ? B : c);
        d[0] = q.g[s].Do; d[1] = q.g[s].Ho; d[2] = q.g[s].Wo;
    }
    return q;
}
struct Carve {
    unsigned char *base; size_t off;
    template <typename T> T *take(size_t n) {
        T *p = reinterpret_cast<T *>(base + off);
        off = align_up(off + n * sizeof(T), 256);
        return p;
    }
};
struct CtSaved { float *h[3]; bf16 *a[2]; float *mu[3], *rstd[3]; bf16 *wr[3]; };
CtSaved carve_ct_saved(void *buf, const CtGeom &q, size_t *bytes) {
    Carve c{static_cast<unsigned char *>(buf), 0};
    CtSaved v;
    for (int s = 0; s < 3; ++s) {
        v.h[s] = c.take<float>((size_t)q.R[s] * q.Cout[s]);
        if (s < 2) v.a[s] = c.take<bf16>((size_t)q.R[s] * q.Cout[s]);
        v.mu[s] = c.take<float>(q.Cout[s]); v.rstd[s] = c.take<float>(q.Cout[s]);
        v.wr[s] = s ? c.take<bf16>((size_t)q.Cout[s] * q.K[s]) : nullptr;
    }
    *bytes = c.off;
    return v;
}
struct CtScratch { void *ctws; bf16 *col, *dcol, *dx; float *dA_a, *dA_b, *dwr; };
CtScratch carve_ct_scratch(void *buf, const CtGeom &q, size_t *bytes) {
    Carve c{static_cast<unsigned char *>(buf), 0};
    CtScratch w;
    w.ctws = c.take<unsigned char>(b200surv_ct_workspace_bytes());
    size_t col = 0, dx = 0, dwr = 0;
    for (int s = 0; s < 3; ++s) {
        if (s) { col = max(col, (size_t)q.chunk[s] * q.vox[s] * q.K[s]); dwr = max(dwr, (size_t)32 * q.Cout[s] * q.K[s]); }
        dx = max(dx, (size_t)q.R[s] * q.Cout[s]);
    }
    w.col = c.take<bf16>(col); w.dcol = c.take<bf16>(col); w.dx = c.take<bf16>(dx);
    w.dA_a = c.take<float>(max((size_t)q.R[2] * q.Cout[2], (size_t)q.R[0] * q.Cout[0]));
    w.dA_b = c.take<float>((size_t)q.R[1] * q.Cout[1]);
    w.dwr = c.take<float>(dwr);
    *bytes = c.off;
    return w;
}
#define CT_TRY(expr) do { const int32_t _rc = (expr); if (_rc != B200SURV_OK) return _rc; } while (0)
}  // namespace
extern "C" {

size_t b200surv_ct_encoder_saved_bytes(int64_t B, int32_t D, int32_t H, int32_t W) {
    if (B < 1 || D < 1 || H < 1 || W < 1) return 0;
    size_t n; carve_ct_saved(nullptr, ct_geom(B, D, H, W), &n); return n;
}
size_t b200surv_ct_encoder_workspace_bytes(int64_t B, int32_t D, int32_t H, int32_t W) {
    if (B < 1 || D < 1 || H < 1 || W < 1) return 0;
    size_t n; carve_ct_scratch(nullptr, ct_geom(B, D, H, W), &n); return n;
}

int32_t b200surv_ct_encoder_fwd(const float *ct, const b200surv_ct_params *p, int64_t B, int32_t D, int32_t H, int32_t W,
                                int32_t training, float *feat, void *saved, size_t saved_bytes, void *workspace,
                                size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(ct && p && feat && saved && workspace, "null pointer");
    B200_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "shape");
    for (int s = 0; s < 3; ++s)
        B200_REQUIRE(p->w[s] && p->gamma[s] && p->beta[s] && p->run_mean[s] && p->run_var[s], "null parameter");
    const CtGeom q = ct_geom(B, D, H, W);
    B200_REQUIRE(q.R[0] * 32 < ((int64_t)1 << 40) && q.chunk[1] * q.vox[1] < ((int64_t)1 << 31), "volume too large");
    size_t need;
    const CtSaved v = carve_ct_saved(saved, q, &need);
    if (saved_bytes < need) { set_error("ct encoder: saved buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    const CtScratch w = carve_ct_scratch(workspace, q, &need);
    if (workspace_bytes < need) { set_error("ct encoder: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    const size_t ctws = b200surv_ct_workspace_bytes();
    for (int s = 0; s < 3; ++s) {
        const Grid3 &g = q.g[s];
        if (s == 0) {
            CT_TRY(b200surv_ct_conv_first_fwd(ct, p->w[0], p->b[0], B, g.D, g.H, g.W, q.Cout[0], v.h[0], stream));
        } else {
            CT_TRY(b200surv_ct_weight_pack(p->w[s], q.Cout[s], q.Cin[s], v.wr[s], stream));
            for (int64_t b0 = 0; b0 < B; b0 += q.chunk[s]) {
                const int64_t bc = min(q.chunk[s], B - b0);
                CT_TRY(b200surv_ct_im2col(v.a[s - 1] + b0 * q.vin[s] * q.Cin[s], bc, g.D, g.H, g.W, q.Cin[s], w.col, stream));
                CT_TRY(gemm_bf16(w.col, q.K[s], 0, v.wr[s], q.K[s], 0, (int)(bc * q.vox[s]), q.Cout[s], q.K[s],
                                 v.h[s] + b0 * q.vox[s] * q.Cout[s], q.Cout[s], nullptr, 0, p->b[s], 0, nullptr, st));
            }
        }
        CT_TRY(b200surv_ct_bn_stats(v.h[s], q.R[s], q.Cout[s], training, p->run_mean[s], p->run_var[s], v.mu[s], v.rstd[s],
                                    w.ctws, ctws, stream));
        if (s < 2)
            CT_TRY(b200surv_ct_bn_relu(v.h[s], v.mu[s], v.rstd[s], p->gamma[s], p->beta[s], q.R[s], q.Cout[s], v.a[s], stream));
        else
            CT_TRY(b200surv_ct_bn_relu_pool(v.h[s], v.mu[s], v.rstd[s], p->gamma[s], p->beta[s], B, (int)q.vox[s], q.Cout[s],
                                            feat, stream));
    }
    return B200SURV_OK;
}

int32_t b200surv_ct_encoder_bwd(const float *ct, const b200surv_ct_params *p, const float *d_feat, int64_t B, int32_t D,
                                int32_t H, int32_t W, int32_t training, const b200surv_ct_grads *grads, const void *saved,
                                size_t saved_bytes, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(ct && p && d_feat && grads && saved && workspace, "null pointer");
    B200_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "shape");
    for (int s = 0; s < 3; ++s)
        B200_REQUIRE(p->gamma[s] && p->beta[s] && grads->w[s] && grads->b[s] && grads->gamma[s] && grads->beta[s], "null parameter");
    const CtGeom q = ct_geom(B, D, H, W);
    size_t need;
    const CtSaved v = carve_ct_saved(const_cast<void *>(saved), q, &need);
    if (saved_bytes < need) { set_error("ct encoder: saved buffer too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    const CtScratch w = carve_ct_scratch(workspace, q, &need);
    if (workspace_bytes < need) { set_error("ct encoder: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    const size_t ctws = b200surv_ct_workspace_bytes();
    float *dA = w.dA_a;
    CT_TRY(b200surv_ct_pool_bwd(d_feat, B, (int)q.vox[2], q.Cout[2], dA, stream));
    for (int s = 2; s >= 0; --s) {
        const Grid3 &g = q.g[s];
        CT_TRY(b200surv_ct_bn_bwd(v.h[s], dA, v.mu[s], v.rstd[s], p->gamma[s], p->beta[s], q.R[s], q.Cout[s], training, w.dx,
                                  grads->gamma[s], grads->beta[s], grads->b[s], w.ctws, ctws, stream));
        if (s == 0) {
            CT_TRY(b200surv_ct_conv_first_wgrad(ct, w.dx, B, g.D, g.H, g.W, q.Cout[0], grads->w[0], w.ctws, ctws, stream));
            break;
        }
        float *dA_prev = (s == 2) ? w.dA_b : w.dA_a;
        for (int64_t b0 = 0; b0 < B; b0 += q.chunk[s]) {
            const int64_t bc = min(q.chunk[s], B - b0);
            const int rows = (int)(bc * q.vox[s]);
            const bf16 *dxc = w.dx + b0 * q.vox[s] * q.Cout[s];
            CT_TRY(b200surv_ct_im2col(v.a[s - 1] + b0 * q.vin[s] * q.Cin[s], bc, g.D, g.H, g.W, q.Cin[s], w.col, stream));
            // dW (tap-major) = dx^T col, both operands as stored (MN-major), split along K over the SMs
            const int nsl = splitk_slices(q.Cout[s], q.K[s], rows, nullptr);
            CT_TRY(gemm_bf16(dxc, q.Cout[s], 1, w.col, q.K[s], 1, q.Cout[s], q.K[s], rows, w.dwr, q.K[s], nullptr, 0, nullptr, 0,
                             w.dwr, st));
            CT_TRY(b200surv_ct_weight_unpack(w.dwr, nsl, q.Cout[s], q.Cin[s], b0 > 0, grads->w[s], stream));
            // dcol = dx W; every input voxel then gathers its taps
            CT_TRY(gemm_bf16(dxc, q.Cout[s], 0, v.wr[s], q.K[s], 1, rows, q.K[s], q.Cout[s], nullptr, 0, w.dcol, q.K[s], nullptr,
                             0, nullptr, st));
            CT_TRY(b200surv_ct_col2im(w.dcol, bc, g.D, g.H, g.W, q.Cin[s], dA_prev + b0 * q.vin[s] * q.Cin[s], stream));
        }
        dA = dA_prev;
    }
    return B200SURV_OK;
}

}  // extern "C"
