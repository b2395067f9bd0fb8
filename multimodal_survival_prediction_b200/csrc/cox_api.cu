// extern "C" entry points for the Cox loss and the C-index (declared in include/b200surv.h).
#include <mutex>

#include "common.cuh"

namespace b200surv {
// cox_binned.cu
size_t cox_binned_state_bytes(int64_t n_seg, int nb);
size_t cox_binned_workspace_bytes(int64_t n, int64_t n_seg, int nb);
int32_t cox_binned_partial(const float *, const float *, const uint8_t *, const int64_t *, int64_t, int64_t, int,
                           float, int64_t *, float *, void *, size_t, cudaStream_t);
int32_t cox_binned_finalize(const int64_t *, const float *, int64_t, int64_t, int, int, int, float, float *, void *,
                            size_t, void *, size_t, cudaStream_t);
int32_t cox_binned_fwd(const float *, const float *, const uint8_t *, const int64_t *, int64_t, int64_t, int, int,
                       int, float, float *, void *, size_t, void *, size_t, cudaStream_t);
int32_t cox_binned_bwd_launch(const float *, const void *, size_t, const float *, const float *, const uint8_t *,
                              const int64_t *, int64_t, int64_t, int, float *, cudaStream_t);
size_t cox_binned_peer_buffer_bytes(int nb);
size_t cox_binned_peer_trace_offset(int64_t n, int nb);
int32_t cox_binned_fwd_peer(const float *, const float *, const uint8_t *, int64_t, int, int, int, float, float *,
                            void *, size_t, void *, size_t, void *const *, int, int, unsigned, cudaStream_t);
// cox_small.cu
int32_t cox_small_fwd_launch(const float *, const float *, const uint8_t *, const int64_t *, int64_t, int64_t, int,
                             int, float *, void *, size_t, cudaStream_t);
int32_t cox_scale_grad_launch(const float *, const void *, const int64_t *, int64_t, int64_t, float *, cudaStream_t);
// cox_sorted.cu
size_t cox_sorted_workspace_bytes(int64_t n, int64_t n_seg);
int32_t cox_sorted_fwd_launch(const float *, const float *, const uint8_t *, const int64_t *, int64_t, int64_t, int, int, float *,
                              void *, size_t, void *, size_t, cudaStream_t);
int32_t cox_sorted_shard_keys(const float *, const float *, const uint8_t *, int64_t, void *, void *, size_t, cudaStream_t);
int32_t cox_sorted_shard_sort(int64_t, void *, size_t, cudaStream_t);
int32_t cox_sorted_shard_reduce(const float *, int64_t, const void *, int, int, void *, void *, size_t, cudaStream_t);
int32_t cox_sorted_shard_terms(int64_t, int, const void *, int, int, void *, void *, size_t, cudaStream_t);
int32_t cox_sorted_shard_finish(int64_t, int, int, const void *, int, int, float *, void *, size_t, void *, size_t, cudaStream_t);
// cindex.cu
size_t cindex_workspace_bytes(int64_t n, int algo);
int32_t cindex_counts_launch(const float *, const float *, const uint8_t *, int64_t, int64_t, int64_t, float, int, int,
                             int, int64_t *, void *, size_t, cudaStream_t);
size_t debug_sortscan_temp_bytes(int64_t n);
int32_t debug_sort_pairs(uint32_t *, uint32_t *, uint32_t *, uint32_t *, int64_t, void *, cudaStream_t);
int32_t debug_scan(const double *, const long long *, int64_t, int, int, double *, double *, long long *, void *, cudaStream_t);
}  // namespace b200surv

using namespace b200surv;

namespace {
// internal streams of b200surv_cindex_counts_cohorts, created once per device on first use
constexpr int COHORT_LANES = 8;
struct CohortLanes {
    cudaStream_t s[COHORT_LANES];
    cudaEvent_t done[COHORT_LANES], fork;
};
CohortLanes *cohort_lanes() {
    static CohortLanes pool[64];
    static int state[64];  // 0 = not tried, 1 = ready, -1 = failed
    static std::mutex mu;  // lazily built once per device; two host threads may arrive together
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (state[dev] == 0) {
        bool ok = cudaEventCreateWithFlags(&pool[dev].fork, cudaEventDisableTiming) == cudaSuccess;
        for (int j = 0; ok && j < COHORT_LANES; ++j)
            ok = cudaStreamCreateWithFlags(&pool[dev].s[j], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&pool[dev].done[j], cudaEventDisableTiming) == cudaSuccess;
        state[dev] = ok ? 1 : -1;
    }
    return state[dev] == 1 ? &pool[dev] : nullptr;
}
}  // namespace

extern "C" {

size_t b200surv_cox_state_bytes(int64_t n, int64_t n_seg, int32_t mode, int32_t nbins) {
    if (n < 0 || n_seg < 1) return 0;
    if (mode == B200SURV_COX_BINNED) return cox_binned_state_bytes(n_seg, nbins);
    return (size_t)n_seg * sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
}

size_t b200surv_cox_workspace_bytes(int64_t n, int64_t n_seg, int32_t mode, int32_t nbins) {
    if (n < 0 || n_seg < 1) return 0;
    if (mode == B200SURV_COX_BINNED) return cox_binned_workspace_bytes(n, n_seg, nbins);
    if (mode == B200SURV_COX_SORTED) return cox_sorted_workspace_bytes(n, n_seg);
    return 256;
}

size_t b200surv_cox_bins_sum_count(int32_t nbins) { return 3 * (size_t)nbins + 4; }

int32_t b200surv_cox_fwd(const float *log_hz, const float *time, const uint8_t *event,
                         const int64_t *seg_offsets, int64_t n, int64_t n_seg, int32_t ties,
                         int32_t reduction, int32_t mode, int32_t nbins, float shift, float *out_loss,
                         void *state, size_t state_bytes, void *workspace, size_t workspace_bytes,
                         b200surv_stream_t stream) {
    B200_REQUIRE(log_hz && time && event && out_loss && state, "null pointer");
    B200_REQUIRE(n >= 1, "n must be >= 1");
    B200_REQUIRE(n_seg >= 1, "n_seg must be >= 1");
    B200_REQUIRE(n_seg == 1 || seg_offsets != nullptr, "seg_offsets required when n_seg > 1");
    cudaStream_t st = as_stream(stream);
    switch (mode) {
        case B200SURV_COX_SMALL:
            return cox_small_fwd_launch(log_hz, time, event, seg_offsets, n, n_seg, ties, reduction, out_loss, state,
                                        state_bytes, st);
        case B200SURV_COX_BINNED:
            B200_REQUIRE(workspace != nullptr, "workspace");
            return cox_binned_fwd(log_hz, time, event, seg_offsets, n, n_seg, ties, reduction, nbins, shift, out_loss,
                                  state, state_bytes, workspace, workspace_bytes, st);
        case B200SURV_COX_SORTED:
            B200_REQUIRE(workspace != nullptr, "workspace");
            return cox_sorted_fwd_launch(log_hz, time, event, seg_offsets, n, n_seg, ties, reduction, out_loss, state, state_bytes,
                                         workspace, workspace_bytes, st);
        default:
            set_error("unknown cox mode %d", mode);
            return B200SURV_BAD_ARG;
    }
}

int32_t b200surv_cox_bwd(const float *grad_out, const void *state, size_t state_bytes, const float *log_hz,
                         const float *time, const uint8_t *event, const int64_t *seg_offsets, int64_t n,
                         int64_t n_seg, int32_t mode, int32_t nbins, float *out_grad, b200surv_stream_t stream) {
    B200_REQUIRE(grad_out && state && out_grad, "null pointer");
    B200_REQUIRE(n >= 1 && n_seg >= 1, "n, n_seg");
    B200_REQUIRE(n_seg == 1 || seg_offsets != nullptr, "seg_offsets required when n_seg > 1");
    cudaStream_t st = as_stream(stream);
    if (mode == B200SURV_COX_BINNED) {
        B200_REQUIRE(log_hz && time && event, "BINNED backward re-reads the inputs");
        return cox_binned_bwd_launch(grad_out, state, state_bytes, log_hz, time, event, seg_offsets, n, n_seg, nbins,
                                     out_grad, st);
    }
    if (mode == B200SURV_COX_SMALL || mode == B200SURV_COX_SORTED) {
        const size_t need = (size_t)n_seg * sizeof(b200surv_cox_header) + (size_t)n * sizeof(float);
        if (state_bytes < need) { set_error("cox bwd: state buffer %zu < %zu", state_bytes, need); return B200SURV_WORKSPACE_TOO_SMALL; }
        return cox_scale_grad_launch(grad_out, state, seg_offsets, n, n_seg, out_grad, st);
    }
    set_error("unknown cox mode %d", mode);
    return B200SURV_BAD_ARG;
}

int32_t b200surv_cox_binned_partial(const float *log_hz, const float *time, const uint8_t *event,
                                    const int64_t *seg_offsets, int64_t n, int64_t n_seg, int32_t nbins,
                                    float shift, int64_t *bins_sum, float *bins_max, void *workspace,
                                    size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(log_hz && time && event && bins_sum && bins_max && workspace, "null pointer");
    B200_REQUIRE(n >= 1 && n_seg >= 1, "n, n_seg");
    B200_REQUIRE(n_seg == 1 || seg_offsets != nullptr, "seg_offsets required when n_seg > 1");
    return cox_binned_partial(log_hz, time, event, seg_offsets, n, n_seg, nbins, shift, bins_sum, bins_max, workspace,
                              workspace_bytes, as_stream(stream));
}

int32_t b200surv_cox_binned_finalize(const int64_t *bins_sum, const float *bins_max, int64_t n, int64_t n_seg,
                                     int32_t ties, int32_t reduction, int32_t nbins, float shift, float *out_loss,
                                     void *state, size_t state_bytes, void *workspace, size_t workspace_bytes,
                                     b200surv_stream_t stream) {
    B200_REQUIRE(bins_sum && bins_max && out_loss && state && workspace, "null pointer");
    B200_REQUIRE(n >= 1 && n_seg >= 1, "n, n_seg");
    return cox_binned_finalize(bins_sum, bins_max, n, n_seg, ties, reduction, nbins, shift, out_loss, state,
                               state_bytes, workspace, workspace_bytes, as_stream(stream));
}

size_t b200surv_cox_peer_buffer_bytes(int32_t nbins) { return cox_binned_peer_buffer_bytes(nbins); }
size_t b200surv_cox_peer_trace_offset(int64_t n, int32_t nbins) { return cox_binned_peer_trace_offset(n, nbins); }

int32_t b200surv_cox_binned_fwd_peer(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                                     int32_t ties, int32_t reduction, int32_t nbins, float shift, float *out_loss,
                                     void *state, size_t state_bytes, void *workspace, size_t workspace_bytes,
                                     void *const *peer_bufs, int32_t world, int32_t rank, uint32_t epoch,
                                     b200surv_stream_t stream) {
    B200_REQUIRE(log_hz && time && event && out_loss && state && workspace && peer_bufs, "null pointer");
    B200_REQUIRE(n >= 1, "n must be >= 1");
    return cox_binned_fwd_peer(log_hz, time, event, n, ties, reduction, nbins, shift, out_loss, state, state_bytes,
                               workspace, workspace_bytes, peer_bufs, world, rank, epoch, as_stream(stream));
}

size_t b200surv_cox_shard_record_bytes(void) { return 128; }

int32_t b200surv_cox_sorted_shard_keys(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                                       void *rec0_out, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(log_hz && time && event && rec0_out, "null pointer");
    return cox_sorted_shard_keys(log_hz, time, event, n, rec0_out, workspace, workspace_bytes, as_stream(stream));
}
int32_t b200surv_cox_sorted_shard_sort(int64_t n, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    return cox_sorted_shard_sort(n, workspace, workspace_bytes, as_stream(stream));
}
int32_t b200surv_cox_sorted_shard_reduce(const float *log_hz, int64_t n, const void *all_rec0, int32_t rank, int32_t world,
                                         void *rec1_out, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(log_hz && all_rec0 && rec1_out, "null pointer");
    return cox_sorted_shard_reduce(log_hz, n, all_rec0, rank, world, rec1_out, workspace, workspace_bytes, as_stream(stream));
}
int32_t b200surv_cox_sorted_shard_terms(int64_t n, int32_t ties, const void *all_rec1, int32_t rank, int32_t world,
                                        void *rec2_out, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(all_rec1 && rec2_out, "null pointer");
    return cox_sorted_shard_terms(n, ties, all_rec1, rank, world, rec2_out, workspace, workspace_bytes, as_stream(stream));
}
int32_t b200surv_cox_sorted_shard_finish(int64_t n, int32_t ties, int32_t reduction, const void *all_rec2, int32_t rank,
                                         int32_t world, float *out_loss, void *state, size_t state_bytes, void *workspace,
                                         size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(all_rec2 && out_loss && state, "null pointer");
    return cox_sorted_shard_finish(n, ties, reduction, all_rec2, rank, world, out_loss, state, state_bytes, workspace,
                                   workspace_bytes, as_stream(stream));
}

size_t b200surv_cindex_workspace_bytes(int64_t n, int64_t n_seg, int32_t algo) {
    if (n < 0 || n_seg < 1) return 0;
    return cindex_workspace_bytes(n, algo);
}

int32_t b200surv_cindex_counts(const float *estimate, const float *time, const uint8_t *event,
                               const int64_t *seg_offsets, int64_t n, int64_t n_seg, int64_t row_begin,
                               int64_t row_end, float tied_tol, int32_t algo, int64_t *out_counts, void *workspace,
                               size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(estimate && time && event && out_counts, "null pointer");
    B200_REQUIRE(n >= 0, "n");
    if (n_seg != 1 || seg_offsets != nullptr) {
        set_error("segmented C-index: call once per cohort (n_seg must be 1 in this version)");
        return B200SURV_UNSUPPORTED;
    }
    B200_REQUIRE(algo == 0 || workspace != nullptr, "workspace");
    return cindex_counts_launch(estimate, time, event, n, row_begin, row_end, tied_tol, algo, 0, 1, out_counts, workspace,
                                workspace_bytes, as_stream(stream));
}

int32_t b200surv_cindex_counts_shard(const float *estimate, const float *time, const uint8_t *event, int64_t n,
                                     int32_t shard, int32_t n_shards, float tied_tol, int64_t *out_counts,
                                     void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(estimate && time && event && out_counts && workspace, "null pointer");
    B200_REQUIRE(n >= 0, "n");
    return cindex_counts_launch(estimate, time, event, n, 0, n, tied_tol, 1, shard, n_shards, out_counts, workspace,
                                workspace_bytes, as_stream(stream));
}

int32_t b200surv_cindex_counts_shard_algo(const float *estimate, const float *time, const uint8_t *event, int64_t n,
                                          int32_t shard, int32_t n_shards, float tied_tol, int32_t algo,
                                          int64_t *out_counts, void *workspace, size_t workspace_bytes,
                                          b200surv_stream_t stream) {
    B200_REQUIRE(estimate && time && event && out_counts && workspace, "null pointer");
    B200_REQUIRE(n >= 0, "n");
    B200_REQUIRE(algo == 1 || algo == 2, "tile shards need algo 1 or 2");
    return cindex_counts_launch(estimate, time, event, n, 0, n, tied_tol, algo, shard, n_shards, out_counts, workspace,
                                workspace_bytes, as_stream(stream));
}

int32_t b200surv_cindex_counts_cohorts(const float *estimate, const float *time, const uint8_t *event,
                                       const int64_t *cohort_offsets_host, int64_t n_cohorts, float tied_tol,
                                       int32_t algo, int64_t *out_counts, void *workspace, size_t workspace_bytes,
                                       b200surv_stream_t stream) {
    B200_REQUIRE(estimate && time && event && out_counts && cohort_offsets_host, "null pointer");
    B200_REQUIRE(n_cohorts >= 1, "n_cohorts");
    B200_REQUIRE(algo == 0 || workspace != nullptr, "workspace");
    cudaStream_t user = as_stream(stream);
    // A cohort is a chain of ~25 small launches; with a workspace that holds k >= 2 cohorts' scratch (k x
    // b200surv_cindex_workspace_bytes(n_max, 1, algo), up to 8) the chains run k at a time on internal streams that
    // fork from and join the caller's stream with events (so everything stays ordered on `stream`).
    int64_t n_max = 0;
    for (int64_t c = 0; c < n_cohorts; ++c) {
        const int64_t a = cohort_offsets_host[c], b = cohort_offsets_host[c + 1];
        B200_REQUIRE(a >= 0 && b >= a, "cohort_offsets_host must be non-decreasing");
        if (b - a > n_max) n_max = b - a;
    }
    const size_t ws_one = align_up(cindex_workspace_bytes(n_max, algo), 256);
    int k = (algo == 0 || ws_one == 0) ? 1 : (int)(workspace_bytes / ws_one);
    if (k > COHORT_LANES) k = COHORT_LANES;
    if (k > n_cohorts) k = (int)n_cohorts;
    CohortLanes *lanes = k >= 2 ? cohort_lanes() : nullptr;
    if (lanes == nullptr) k = 1;
    if (k >= 2) {
        B200_CHECK_CUDA(cudaEventRecord(lanes->fork, user));
        for (int j = 0; j < k; ++j) B200_CHECK_CUDA(cudaStreamWaitEvent(lanes->s[j], lanes->fork, 0));
    }
    for (int64_t c = 0; c < n_cohorts; ++c) {
        const int64_t a = cohort_offsets_host[c], b = cohort_offsets_host[c + 1];
        if (b == a) continue;
        const int j = (int)(c % k);
        const int32_t rc = cindex_counts_launch(estimate + a, time + a, event + a, b - a, 0, b - a, tied_tol, algo, 0, 1,
                                                out_counts + 6 * c,
                                                k >= 2 ? static_cast<unsigned char *>(workspace) + (size_t)j * ws_one : workspace,
                                                k >= 2 ? ws_one : workspace_bytes, k >= 2 ? lanes->s[j] : user);
        if (rc != B200SURV_OK) return rc;
    }
    if (k >= 2) {
        for (int j = 0; j < k; ++j) {
            B200_CHECK_CUDA(cudaEventRecord(lanes->done[j], lanes->s[j]));
            B200_CHECK_CUDA(cudaStreamWaitEvent(user, lanes->done[j], 0));
        }
    }
    return B200SURV_OK;
}

/* test hooks for the sort / scan primitives (csrc/sortscan.cuh) */
size_t b200surv_debug_sortscan_temp_bytes(int64_t n) { return debug_sortscan_temp_bytes(n); }
int32_t b200surv_debug_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                                  void *temp, b200surv_stream_t stream) {
    B200_REQUIRE(keys && vals && keys_tmp && vals_tmp && temp && n >= 1, "arguments");
    return debug_sort_pairs(keys, vals, keys_tmp, vals_tmp, n, temp, as_stream(stream));
}
int32_t b200surv_debug_scan(const double *a, const int64_t *i, int64_t n, int32_t iop, int32_t reverse, double *out_a,
                            double *out_b, int64_t *out_i, void *temp, b200surv_stream_t stream) {
    B200_REQUIRE(a && i && out_a && out_b && out_i && temp && n >= 1 && iop >= 0 && iop <= 2, "arguments");
    return debug_scan(a, reinterpret_cast<const long long *>(i), n, iop, reverse, out_a, out_b,
                      reinterpret_cast<long long *>(out_i), temp, as_stream(stream));
}

}  // extern "C"
