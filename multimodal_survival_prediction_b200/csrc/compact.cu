// Labelled-row compaction.
//
// The reference's training step keeps only the rows that carry a survival label before the loss
// (scripts/training/partial_modality_training.py:401-408; simple_fusion.py:255-268):
//     survival_mask = torch.tensor(has_survival, dtype=torch.bool);  hazard[mask], label[mask, 0], label[mask, 1]
// and skips the loss unless n >= 2 rows are left and at least one is an event.  Here: one exclusive prefix sum of the
// flags (the single-pass look-back scan of sortscan.cuh) scatters hazard / time / event of the kept rows, order
// preserved, and leaves {kept rows, kept events} on the device; the backward pass scatters the loss gradient back.
#include "common.cuh"
#include "sortscan.cuh"

namespace b200surv {
namespace {

struct LoadKeep {
    const uint8_t *keep; const float *label;
    __device__ sortscan::Tup operator()(int64_t p) const {
        sortscan::Tup t;
        const bool k = keep[p] != 0;
        t.a = (k && label[2 * p + 1] != 0.f) ? 1.0 : 0.0;  // kept events
        t.b = 0.0; t.i = k ? 1 : 0;
        return t;
    }
};
struct StoreCompact {
    const float *hazard, *label; int64_t B;
    float *out_hazard, *out_time; uint8_t *out_event; int *out_index; long long *out_counts;
    __device__ void operator()(int64_t p, const sortscan::Tup &inc, const sortscan::Tup &el) const {
        if (el.i) {
            const long long q = inc.i - 1;
            out_hazard[q] = hazard[p]; out_time[q] = label[2 * p]; out_event[q] = label[2 * p + 1] != 0.f ? 1 : 0;
            out_index[q] = (int)p;
        }
        if (p == B - 1) { out_counts[0] = inc.i; out_counts[1] = (long long)inc.a; }
    }
};

__global__ void k_scatter_rows(const float *__restrict__ grad_sel, const int *__restrict__ index, int64_t n_sel,
                               float *__restrict__ out_grad) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_sel; q += (int64_t)gridDim.x * blockDim.x)
        out_grad[index[q]] = grad_sel[q];
}

}  // namespace
}  // namespace b200surv

using namespace b200surv;

extern "C" {

size_t b200surv_compact_workspace_bytes(int64_t B) { return sortscan::scan_state_bytes(B > 0 ? B : 1) + 256; }

int32_t b200surv_compact_labelled(const float *hazard, const float *label, const uint8_t *has_survival, int64_t B,
                                  float *out_hazard, float *out_time, uint8_t *out_event, int32_t *out_index,
                                  int64_t *out_counts, void *workspace, size_t workspace_bytes, b200surv_stream_t stream) {
    B200_REQUIRE(hazard && label && has_survival && out_hazard && out_time && out_event && out_index && out_counts && workspace,
                 "null pointer");
    B200_REQUIRE(B >= 1 && B < ((int64_t)1 << 31), "B in [1, 2^31)");
    if (workspace_bytes < b200surv_compact_workspace_bytes(B)) { set_error("compact: workspace too small"); return B200SURV_WORKSPACE_TOO_SMALL; }
    return sortscan::scan_lookback<sortscan::I_ADD, false>(
        B, LoadKeep{has_survival, label},
        StoreCompact{hazard, label, B, out_hazard, out_time, out_event, out_index, reinterpret_cast<long long *>(out_counts)},
        workspace, as_stream(stream));
}

int32_t b200surv_scatter_rows(const float *grad_sel, const int32_t *index, int64_t n_sel, int64_t B, float *out_grad,
                              b200surv_stream_t stream) {
    B200_REQUIRE(out_grad && B >= 1 && n_sel >= 0 && n_sel <= B, "arguments");
    cudaStream_t st = as_stream(stream);
    B200_CHECK_CUDA(cudaMemsetAsync(out_grad, 0, (size_t)B * sizeof(float), st));
    if (n_sel > 0) {
        B200_REQUIRE(grad_sel && index, "null pointer");
        int grid = (int)((n_sel + 255) / 256);
        if (grid > 8 * num_sms()) grid = 8 * num_sms();
        k_scatter_rows<<<grid, 256, 0, st>>>(grad_sel, index, n_sel, out_grad);
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return B200SURV_OK;
}

}  // extern "C"
