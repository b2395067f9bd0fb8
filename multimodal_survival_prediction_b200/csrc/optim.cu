// Fused clip_grad_norm_(max_norm) + Adam / AdamW step over a list of fp32 parameter tensors.
//
// The reference's training step ends with torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) followed by
// optimizer.step() with optim.Adam(lr, weight_decay=1e-4) (scripts/training/partial_modality_training.py:427-428,536;
// final_multimodal.py:259,350) or optim.AdamW (simple_fusion.py:273-274,391).  Two passes over the gradients instead of
// the framework's per-tensor kernels: (1) sum of squares per 4096-element unit, summed in a fixed order
// (deterministic) -> total norm -> clip coefficient; (2) the update, reading grad, m, v, p once and writing m, v, p.
#include "common.cuh"

namespace b200surv {
namespace {

constexpr int OPT_MAX_TENSORS = 64;   // per launch (the list travels in the kernel parameters)
constexpr int OPT_UNIT = 4096;        // elements per work unit
constexpr int OPT_THREADS = 256;

struct OptList {
    float *p[OPT_MAX_TENSORS], *g[OPT_MAX_TENSORS], *m[OPT_MAX_TENSORS], *v[OPT_MAX_TENSORS];
    long long unit0[OPT_MAX_TENSORS + 1];  // first work unit of each tensor (exclusive prefix of ceil(numel / OPT_UNIT))
    long long numel[OPT_MAX_TENSORS];
    int n;
};

__device__ __forceinline__ int tensor_of(const OptList &L, long long unit) {
    int lo = 0, hi = L.n;  // unit0[lo] <= unit < unit0[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (L.unit0[mid] <= unit) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(OPT_THREADS)
k_sumsq(const __grid_constant__ OptList L, long long unit_base, double *__restrict__ partial) {
    __shared__ double red[32];
    const long long unit = blockIdx.x;
    const int t = tensor_of(L, unit);
    const long long e0 = (unit - L.unit0[t]) * OPT_UNIT, e1 = min(L.numel[t], e0 + OPT_UNIT);
    const float *g = L.g[t];
    double s = 0.0;
    for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) { const float x = g[i]; s += (double)x * (double)x; }
    s = block_reduce<double>(s, 0.0, OpAddD(), red);
    if (threadIdx.x == 0) partial[unit_base + unit] = s;
}

// total norm (fixed order), clip coefficient like torch: min(1, max_norm / (norm + 1e-6)); max_norm <= 0: no clipping
__global__ void __launch_bounds__(1024)
k_norm(const double *__restrict__ partial, long long n_units, float max_norm, float *__restrict__ coef_norm /*[2]*/) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n_units; i += blockDim.x) s += partial[i];
    s = block_reduce<double>(s, 0.0, OpAddD(), red);
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(s);
        float c = 1.f;
        if (max_norm > 0.f) { c = max_norm / (norm + 1e-6f); if (c > 1.f) c = 1.f; }
        coef_norm[0] = c; coef_norm[1] = norm;
    }
}

__global__ void __launch_bounds__(OPT_THREADS)
k_adam(const __grid_constant__ OptList L, const float *__restrict__ coef_norm, float lr, float beta1, float beta2, float eps,
       float wd, int adamw, float bias1, float bias2_sqrt) {
    const long long unit = blockIdx.x;
    const int t = tensor_of(L, unit);
    const long long e0 = (unit - L.unit0[t]) * OPT_UNIT, e1 = min(L.numel[t], e0 + OPT_UNIT);
    float *p = L.p[t], *m = L.m[t], *v = L.v[t];
    const float *g = L.g[t];
    const float c = coef_norm[0], step_size = lr / bias1;
    for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) {
        float x = p[i], gi = g[i] * c;
        if (adamw) x *= 1.f - lr * wd; else gi = fmaf(wd, x, gi);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi; v[i] = vi;
        p[i] = x - step_size * mi / (sqrtf(vi) / bias2_sqrt + eps);
    }
}

}  // namespace
}  // namespace b200surv

using namespace b200surv;

extern "C" {

size_t b200surv_clip_adam_workspace_bytes(const int64_t *numel_host, int32_t n_tensors) {
    long long units = 0;
    for (int i = 0; i < n_tensors; ++i) units += (numel_host[i] + OPT_UNIT - 1) / OPT_UNIT;
    return align_up((size_t)units * sizeof(double), 256) + 256;
}

int32_t b200surv_clip_adam_step(float *const *params, const float *const *grads, float *const *exp_avg,
                                float *const *exp_avg_sq, const int64_t *numel_host, int32_t n_tensors, float max_norm,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int32_t adamw,
                                int64_t step, float *out_total_norm, void *workspace, size_t workspace_bytes,
                                b200surv_stream_t stream) {
    B200_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel_host && workspace, "null pointer");
    B200_REQUIRE(n_tensors >= 1 && step >= 1, "n_tensors >= 1, step counts from 1");
    if (workspace_bytes < b200surv_clip_adam_workspace_bytes(numel_host, n_tensors)) {
        set_error("clip+adam: workspace too small");
        return B200SURV_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = as_stream(stream);
    long long total_units = 0;
    for (int i = 0; i < n_tensors; ++i) {
        B200_REQUIRE(numel_host[i] >= 1 && params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "tensor list");
        total_units += (numel_host[i] + OPT_UNIT - 1) / OPT_UNIT;
    }
    double *partial = static_cast<double *>(workspace);
    float *coef_norm = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) +
                                                 align_up((size_t)total_units * sizeof(double), 256));
    const float bias1 = 1.f - powf(beta1, (float)step), bias2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    for (int pass = 0; pass < 2; ++pass) {
        long long unit_base = 0;
        for (int t0 = 0; t0 < n_tensors; t0 += OPT_MAX_TENSORS) {
            OptList L;
            L.n = n_tensors - t0 < OPT_MAX_TENSORS ? n_tensors - t0 : OPT_MAX_TENSORS;
            long long u = 0;
            for (int i = 0; i < L.n; ++i) {
                L.p[i] = params[t0 + i]; L.g[i] = const_cast<float *>(grads[t0 + i]); L.m[i] = exp_avg[t0 + i];
                L.v[i] = exp_avg_sq[t0 + i]; L.numel[i] = numel_host[t0 + i]; L.unit0[i] = u;
                u += (numel_host[t0 + i] + OPT_UNIT - 1) / OPT_UNIT;
            }
            L.unit0[L.n] = u;
            if (pass == 0) k_sumsq<<<(unsigned)u, OPT_THREADS, 0, st>>>(L, unit_base, partial);
            else k_adam<<<(unsigned)u, OPT_THREADS, 0, st>>>(L, coef_norm, lr, beta1, beta2, eps, weight_decay, adamw, bias1, bias2_sqrt);
            unit_base += u;
        }
        if (pass == 0) k_norm<<<1, 1024, 0, st>>>(partial, total_units, max_norm, coef_norm);
    }
    if (out_total_norm) B200_CHECK_CUDA(cudaMemcpyAsync(out_total_norm, coef_norm + 1, sizeof(float), cudaMemcpyDeviceToDevice, st));
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // extern "C"
