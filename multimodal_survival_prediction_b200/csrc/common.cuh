// Shared device/host helpers for libb200surv (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200surv.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200surv is written for sm_100a (B200) only"
#endif

namespace b200surv {

void set_error(const char *fmt, ...);

#define B200_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::b200surv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                                  cudaGetErrorString(_e));                                     \
            return B200SURV_CUDA_ERROR;                                                        \
        }                                                                                      \
    } while (0)

#define B200_REQUIRE(cond, msg)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::b200surv::set_error("%s:%d: bad argument: %s (%s)", __FILE__, __LINE__, msg, #cond); \
            return B200SURV_BAD_ARG;                                                           \
        }                                                                                      \
    } while (0)

inline cudaStream_t as_stream(b200surv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();         // of the CURRENT device, cached per device (148 on B200)
int current_device();  // cudaGetDevice, -1 on failure
void count_launches(int k);  // adds to the process-wide counter behind b200surv_debug_launch_count()

// One-time per-DEVICE setup (cudaFuncSetAttribute is a per-device attribute: a process that drives several GPUs must
// set it on each).  Setting an attribute twice is harmless, so two threads racing through `pending()` is fine; the
// flags themselves are atomics, there is no other shared mutable state in the library's launch paths.
constexpr int MAX_DEVICES = 64;
struct PerDeviceOnce {
    volatile unsigned char done[MAX_DEVICES];
    bool pending() const { const int d = current_device(); return d < 0 || d >= MAX_DEVICES || !done[d]; }
    void mark() { const int d = current_device(); if (d >= 0 && d < MAX_DEVICES) done[d] = 1; }
};

// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

// ---- programmatic dependent launch for chains of short kernels (head, CT encoder): a chained kernel lets its successor
// start its own set-up at once (pdl_trigger) and waits for its predecessor's results before it touches global memory
// (pdl_wait: returns when the preceding grid has completed and its writes are visible).  Launched without the attribute
// (B200SURV_PDL=0, or behind a kernel of another library) both instructions are no-ops and the edge is a plain one.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }
bool pdl_enabled();   // api.cu: B200SURV_PDL != "0"
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args &&>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_or(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(FULL, v, o);
    return v;
}

// Block-wide reductions through a caller-provided shared scratch of >= 32 elements.
// Result valid in every thread.  All threads of the block must call.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T identity, Op op, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T r = (lane < nw) ? scratch[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(FULL, r, o));
    return r;
}
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpAddF { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpAddLL { __device__ long long operator()(long long a, long long b) const { return a + b; } };
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpOrU { __device__ unsigned operator()(unsigned a, unsigned b) const { return a | b; } };

// float atomic max for any sign (CAS free): positive floats order as ints, negative as reversed uints
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}

// 128-bit / 32-bit read-only loads with the DEFAULT L2 policy: the cohort is read twice per step (forward, then
// backward in reverse order), so the tail of the first pass should stay L2-resident (an evict-first hint here
// cost 8 us in the backward pass -- measured).  Intrinsics rather than asm volatile so that the compiler may
// hoist them above shared-memory atomics (memory-level parallelism).
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) { return __ldg(reinterpret_cast<const unsigned int *>(p)); }
// ---- mbarrier + bulk-copy (TMA) primitives shared by the streaming kernels and the GEMM
__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbarrier_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarrier_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint32_t bar, uint32_t parity) {
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
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier; size % 16 == 0
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// write-once 128-bit store, evict-first (st.global.cs): the gradient must not displace the inputs in L2
// L2 residency hints for kernels that mix a stream with a random gather / scatter over an array that fits the 126 MB L2:
// the stream is marked evict-first, the gathered / scattered array evict-last, so that the 4-byte random accesses merge in
// L2 instead of costing one 32-byte DRAM sector each (a partly written sector is a read-modify-write in HBM).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ldg_hint_u32(const void *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_hint_f32(const void *p, uint64_t pol) { return __uint_as_float(ldg_hint_u32(p, pol)); }
__device__ __forceinline__ void stg_hint_f32(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_stream_f4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }
#endif  // __CUDACC__

}  // namespace b200surv
