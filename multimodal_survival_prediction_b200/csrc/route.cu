// Row-block shards -> time-range shards (multi-GPU SORTED Cox, SURVEY.md 8e path (B); include/b200surv.h "route").
//
// The reference sorts one cohort on one device (argsort(time), scripts/training/partial_modality_training.py:303-309).  With the
// patients sharded by row block over the GPUs of a box (BASELINE.json north_star), the sort on survival time becomes a
// sample sort: every rank cuts the time axis at the same world - 1 splitters, sends each row to the rank that owns its
// range (one all-to-all per vector, issued by the caller), and the SORTED shard phases (cox_sorted.cu) do the rest.  This file
// is the rank-local part: the destination of every row, the rows grouped by destination (stable: the one-pass radix sort of
// sortscan.cuh on the 8-bit destination), the packed send buffers, and the way back for the gradient.
#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

constexpr int ROUTE_MAX_DEST = 64;

struct RouteLayout {
    size_t off_keys, off_vals, off_keys2, off_vals2, off_tmp, total;
};
RouteLayout route_layout(int64_t n) {
    RouteLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.off_keys = take(N * 4 + 4); L.off_vals = take(N * 4); L.off_keys2 = take(N * 4 + 4); L.off_vals2 = take(N * 4);
    L.off_tmp = take(sortscan::radix_sort_temp_bytes((int64_t)N));
    L.total = o;
    return L;
}

// dest = number of splitters <= t (rows equal to a splitter go right; any rule is fine, tie groups may straddle shards);
// NaN / negative times go to the last rank, where the shard phases flag them.  Per-destination counts: shared atomics per CTA,
// one global atomic per (CTA, destination).
__global__ void __launch_bounds__(256)
k_route_keys(const float *__restrict__ time, int64_t n, const float *__restrict__ splitters, int n_dest, uint32_t *__restrict__ keys,
             long long *__restrict__ counts) {
    __shared__ float s_split[ROUTE_MAX_DEST];
    __shared__ unsigned s_cnt[ROUTE_MAX_DEST];
    if (threadIdx.x < ROUTE_MAX_DEST) {
        s_split[threadIdx.x] = threadIdx.x < n_dest - 1 ? splitters[threadIdx.x] : INFINITY;
        s_cnt[threadIdx.x] = 0u;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float t = time[i];
        int lo = 0, hi = n_dest - 1;               // first splitter > t
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_split[mid] <= t) lo = mid + 1; else hi = mid;
        }
        const int d = (t >= 0.f) ? lo : n_dest - 1;
        keys[i] = (uint32_t)d;
        atomicAdd(&s_cnt[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_dest && s_cnt[threadIdx.x]) atomicAdd(reinterpret_cast<unsigned long long *>(counts + threadIdx.x), (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
k_route_gather(const float *__restrict__ log_hz, const float *__restrict__ time, const uint8_t *__restrict__ event,
               const uint32_t *__restrict__ perm, int64_t n, float *__restrict__ out_lh, float *__restrict__ out_t,
               uint8_t *__restrict__ out_ev, int32_t *__restrict__ out_perm) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint32_t r = perm[j];
        if (out_lh) out_lh[j] = log_hz[r];
        if (out_t) out_t[j] = time[r];
        if (out_ev) out_ev[j] = event[r];
        if (out_perm) out_perm[j] = (int32_t)r;
    }
}

__global__ void __launch_bounds__(256)
k_route_scatter(const float *__restrict__ src, const int32_t *__restrict__ perm, int64_t n, float *__restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) dst[perm[j]] = src[j];
}

int grid_for(int64_t n) {
    int g = (int)((n + 255) / 256);
    const int cap = 16 * num_sms();
    return g > cap ? cap : (g < 1 ? 1 : g);
}

}  // namespace
}  // namespace b200surv

using namespace b200surv;

extern "C" {

size_t b200surv_route_workspace_bytes(int64_t n) { return n < 0 ? 0 : route_layout(n).total; }

int32_t b200surv_route_rows(const float *log_hz, const float *time, const uint8_t *event, int64_t n, const float *splitters,
                            int32_t n_dest, float *out_log_hz, float *out_time, uint8_t *out_event, int32_t *out_perm,
                            int64_t *out_counts, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(time && out_perm && out_counts && workspace, "null pointer");
    B200_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) - 2, "n must be in [1, 2^31)");
    B200_REQUIRE(n_dest >= 1 && n_dest <= ROUTE_MAX_DEST, "n_dest must be in [1, 64]");
    B200_REQUIRE(n_dest == 1 || splitters != nullptr, "splitters");
    B200_REQUIRE((out_log_hz == nullptr || log_hz != nullptr) && (out_event == nullptr || event != nullptr), "sources of the outputs");
    const RouteLayout L = route_layout(n);
    if (workspace_bytes < L.total) { set_error("route: workspace %zu < %zu", workspace_bytes, L.total); return B200SURV_WORKSPACE_TOO_SMALL; }
    cudaStream_t st = as_stream(stream);
    unsigned char *w8 = static_cast<unsigned char *>(workspace);
    uint32_t *keys = reinterpret_cast<uint32_t *>(w8 + L.off_keys), *vals = reinterpret_cast<uint32_t *>(w8 + L.off_vals);
    uint32_t *keys2 = reinterpret_cast<uint32_t *>(w8 + L.off_keys2), *vals2 = reinterpret_cast<uint32_t *>(w8 + L.off_vals2);
    B200_CHECK_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)n_dest * sizeof(int64_t), st));
    k_route_keys<<<grid_for(n), 256, 0, st>>>(time, n, splitters, n_dest, keys, reinterpret_cast<long long *>(out_counts));
    int in_first = 1;
    // one pass of 8 bits on the destination; the values are the row numbers (iota: nothing is read for them)
    int32_t rc = sortscan::radix_sort_pairs2(keys, vals, keys2, vals2, n, 8, nullptr, 0, w8 + L.off_tmp, st, &in_first, true);
    if (rc) return rc;
    const uint32_t *perm = in_first ? vals : vals2;
    k_route_gather<<<grid_for(n), 256, 0, st>>>(log_hz, time, event, perm, n, out_log_hz, out_time, out_event, out_perm);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(5);
    return B200SURV_OK;
}

int32_t b200surv_route_gather(const float *src, const int32_t *perm, int64_t n, float *out, b200surv_stream_t stream) {
    B200_REQUIRE(src && perm && out && n >= 1, "arguments");
    k_route_gather<<<grid_for(n), 256, 0, as_stream(stream)>>>(src, nullptr, nullptr, reinterpret_cast<const uint32_t *>(perm), n, out,
                                                              nullptr, nullptr, nullptr);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(1);
    return B200SURV_OK;
}

int32_t b200surv_route_scatter(const float *src, const int32_t *perm, int64_t n, float *out, b200surv_stream_t stream) {
    B200_REQUIRE(src && perm && out && n >= 1, "arguments");
    k_route_scatter<<<grid_for(n), 256, 0, as_stream(stream)>>>(src, perm, n, out);
    B200_CHECK_CUDA(cudaGetLastError());
    count_launches(1);
    return B200SURV_OK;
}

}  // extern "C"
