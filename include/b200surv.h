/* b200surv.h -- C ABI of libb200surv.so: the B200 (sm_100a) survival-training hot path.
 *
 * Drop-in boundary for baek0203/multimodal_survival_prediction (reference paths below are relative
 * to the reference checkout).  The reference is pure Python; what it binds for this path is
 *   torchsurv.loss.cox.neg_partial_log_likelihood(log_hz, event, time)
 *       scripts/training/partial_modality_training.py:285-288, simple_fusion.py:270,311,
 *       final_multimodal.py:158-162
 *   torchsurv.metrics.cindex.ConcordanceIndex()(estimate, event, time)
 *       partial_modality_training.py:290-294, simple_fusion.py:330-331
 *   PartialModalityNet.forward / MultiModalSurvivalNet.forward (the fusion head)
 *       partial_modality_training.py:234-277, final_multimodal.py:122-150
 * A maintainer binds these entry points with ctypes (INTEGRATION.md shows the stub); the Python
 * host layer in multimodal_survival_prediction_b200/ is exactly such a binding.
 *
 * Conventions for EVERY function:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named *_host;
 *   - the caller owns every buffer (inputs, outputs, state, workspace); the library never allocates
 *     device memory, never synchronises and never throws; work is ordered on `stream` only;
 *   - `event` is one byte per row (torch.bool storage), non-zero = event observed;
 *   - returns B200SURV_OK (0) or a negative b200surv_status; b200surv_last_error() (thread-local)
 *     describes the last failure, including the CUDA error string for B200SURV_CUDA_ERROR;
 *   - stateless and re-entrant: safe from PyTorch's autograd worker threads.
 */
#ifndef B200SURV_H_
#define B200SURV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *b200surv_stream_t; /* == cudaStream_t */

typedef enum {
    B200SURV_OK = 0,
    B200SURV_BAD_ARG = -1,
    B200SURV_WORKSPACE_TOO_SMALL = -2,
    B200SURV_UNSUPPORTED_ARCH = -3,
    B200SURV_CUDA_ERROR = -4,
    B200SURV_UNSUPPORTED = -5
} b200surv_status;

/* ---- library ---------------------------------------------------------------------------- */
int32_t b200surv_version(void);                 /* 1000*major + minor */
int32_t b200surv_arch_check(int32_t device);    /* OK iff the device is compute capability 10.x */
const char *b200surv_last_error(void);
/* Diagnostics: kernels launched so far by the Cox entry points of this process (BINNED, SMALL; SORTED counts its own
 * launches, not those of its sort/scan primitives).  bench.py reports the difference over its timed region. */
uint64_t b200surv_debug_launch_count(void);

/* ---- Cox negative partial log-likelihood ------------------------------------------------- */
/* ties: tie handling of the partial likelihood (replaces torchsurv's ties_method string). */
#define B200SURV_TIES_BRESLOW 1
#define B200SURV_TIES_EFRON 2
/* reduction: MEAN_TERMS averages over the terms the method yields (one per event for Breslow,
 * one per distinct event time for Efron -- torchsurv's "mean" as recollected);
 * MEAN_EVENTS always divides by the number of events (the reference fallback's rule,
 * partial_modality_training.py:309); SUM does not divide. */
#define B200SURV_REDUCE_MEAN_TERMS 0
#define B200SURV_REDUCE_SUM 1
#define B200SURV_REDUCE_MEAN_EVENTS 2
/* mode: which device algorithm runs.
 *   SMALL  : one CTA per segment, any float times, every segment <= B200SURV_COX_SMALL_MAX rows.
 *   BINNED : times must be integer-valued in [0, nbins), nbins <= 8192 (days); no sort: two streaming passes over
 *            (log_hz, time, event) = 22 algorithmic bytes per row for fwd+bwd.  Rows that violate
 *            the precondition raise B200SURV_COXF_NOT_BINNABLE in the header and poison the loss
 *            with NaN; the caller then re-runs with SORTED.
 *   SORTED : any non-negative float times, one cohort or packed cohorts; radix sort on (cohort, time, event) + single-pass
 *            look-back scans whose sums restart per tie group / per cohort, fp64: also the path for log-hazards spread over
 *            more nats than the fixed point of BINNED holds (B200SURV_COXF_LOW_PRECISION). */
#define B200SURV_COX_SMALL 1
#define B200SURV_COX_BINNED 2
#define B200SURV_COX_SORTED 3
#define B200SURV_COX_SMALL_MAX 2048
#define B200SURV_COX_MAX_BINS 8192

/* header flags */
#define B200SURV_COXF_NOT_BINNABLE 1u /* a time was non-integer or outside [0, nbins)            */
#define B200SURV_COXF_EXP_RANGE 2u    /* shift unsuitable for the 36.28 fixed-point weights
                                         (max(log_hz) - shift outside [-8, 20], or sum of weights
                                         >= 2^30): re-run with shift near max_log_hz (header)    */
#define B200SURV_COXF_BAD_TIME 4u     /* NaN or negative time                                     */
#define B200SURV_COXF_PEER_TIMEOUT 8u /* b200surv_cox_binned_fwd_peer: a peer rank never arrived   */
#define B200SURV_COXF_LOW_PRECISION 16u /* BINNED: a row's weight exp(log_hz - shift) is below 2^-16, i.e.
                                         under 2^12 quanta of the 36.28 fixed point (min(log_hz) - shift <
                                         -11.09): a risk set made of such rows would lose precision or
                                         round to zero.  Re-run with a larger shift if the spread allows,
                                         else with SORTED (fp64).  The loss is poisoned with NaN.        */

#define B200SURV_COXF_NOT_PARTITIONED 32u /* b200surv_cox_sorted_shard_*: a shard holds a time above the next shard's
                                         smallest time (the shards are not time ranges).  Loss NaN.          */

/* One 64-byte header per segment in the state buffer (device memory).  SMALL and SORTED: the n_seg headers are
 * contiguous at the start of the buffer.  BINNED: segment s owns the slice [s * stride, (s + 1) * stride) with
 * stride = b200surv_cox_state_bytes(n, 1, BINNED, nbins) = 64 + 8 * nbins; its header is the first 64 bytes of
 * that slice, the (P, F) table follows. */
typedef struct {
    uint32_t flags;        /* B200SURV_COXF_* */
    int32_t mode;
    float loss;            /* same value as out_loss[seg] */
    float scale;           /* d loss / d pll = -1/normaliser (0 when the segment has no event) */
    float shift;           /* exponent shift used: weights are exp(log_hz - shift) */
    float max_log_hz;
    float max_time;
    int32_t nbins;
    int64_t n_events;
    int64_t n_event_times; /* distinct times carrying at least one event */
    double pll;            /* partial log-likelihood before reduction */
    float min_log_hz;      /* BINNED only (else 0) */
    int32_t reserved;
} b200surv_cox_header;

/* Bytes of caller-owned state (kept from fwd to bwd) and scratch workspace. */
size_t b200surv_cox_state_bytes(int64_t n, int64_t n_seg, int32_t mode, int32_t nbins);
size_t b200surv_cox_workspace_bytes(int64_t n, int64_t n_seg, int32_t mode, int32_t nbins);

/* Forward: out_loss[n_seg] (device fp32) and the state buffer (headers + what bwd needs).
 * seg_offsets: device int64[n_seg+1] row offsets of independent cohorts packed back to back, or
 * NULL for one cohort of n rows.  shift: exponent shift (0 is right unless |log_hz| is huge).
 * A segment with no event yields loss 0 and a zero gradient (reference fallback rule,
 * partial_modality_training.py:298-301). */
int32_t b200surv_cox_fwd(const float *log_hz, const float *time, const uint8_t *event,
                         const int64_t *seg_offsets, int64_t n, int64_t n_seg, int32_t ties,
                         int32_t reduction, int32_t mode, int32_t nbins, float shift,
                         float *out_loss, void *state, size_t state_bytes, void *workspace,
                         size_t workspace_bytes, b200surv_stream_t stream);

/* Backward: out_grad[n] = grad_out[seg] * d loss[seg] / d log_hz, original row order. */
int32_t b200surv_cox_bwd(const float *grad_out, const void *state, size_t state_bytes,
                         const float *log_hz, const float *time, const uint8_t *event,
                         const int64_t *seg_offsets, int64_t n, int64_t n_seg, int32_t mode,
                         int32_t nbins, float *out_grad, b200surv_stream_t stream);

/* Row-block sharded BINNED forward for multi-GPU: each rank accumulates its rows'
 * per-bin aggregates, the caller all-reduces them (SUM over bins_sum, MAX over bins_max) with
 * NCCL, then every rank finalises identically and runs b200surv_cox_bwd on its own rows.
 *   bins_sum : int64[n_seg][3*nbins + 4]  = S_censored[nbins], S_event[nbins] (sums of
 *              exp(log_hz - shift) in 36.28 fixed point), m[nbins] (event counts), sum of event
 *              log_hz in 2^-24 fixed point, NOT_BINNABLE count, ceil(sum of weights), BAD_TIME count
 *              -- integers, so the SUM all-reduce is exact and independent of the sharding
 *   bins_max : float[n_seg][2] = max log_hz, max time  -- MAX-reducible
 * n is the rank-local row count in both calls (it sizes the workspace). */
size_t b200surv_cox_bins_sum_count(int32_t nbins);
int32_t b200surv_cox_binned_partial(const float *log_hz, const float *time, const uint8_t *event,
                                    const int64_t *seg_offsets, int64_t n, int64_t n_seg,
                                    int32_t nbins, float shift, int64_t *bins_sum, float *bins_max,
                                    void *workspace, size_t workspace_bytes,
                                    b200surv_stream_t stream);
int32_t b200surv_cox_binned_finalize(const int64_t *bins_sum, const float *bins_max, int64_t n,
                                     int64_t n_seg, int32_t ties, int32_t reduction, int32_t nbins, float shift,
                                     float *out_loss, void *state, size_t state_bytes,
                                     void *workspace, size_t workspace_bytes,
                                     b200surv_stream_t stream);

/* The same sharded forward with the exchange FUSED into the kernel over peer memory (NVLink / NVSwitch): no
 * NCCL call on the data path.  Every rank of the box allocates one peer buffer of
 * b200surv_cox_peer_buffer_bytes(nbins) bytes, zero-filled once, that all ranks can address (CUDA IPC / VMM
 * fabric handles / torch symmetric memory); peer_bufs[r] (HOST array of `world` device pointers, 256-byte
 * aligned) is rank r's buffer as mapped into THIS process.  One cooperative launch per rank: pass 1 -> per-bin
 * int64 sums -> the warp that owns a 32-bin block PUSHES its sums into every peer's buffer (plain remote stores,
 * each 64-bit word tagged with epoch & 3 in its top bits: a word is its own "arrived" mark, so there is no flag, no
 * fence and no extra barrier), polls its own buffer until every source's words carry the tag (2 s time-out ->
 * B200SURV_COXF_PEER_TIMEOUT, loss NaN) and adds them -> the O(nbins) tail, loss, header and (P,F) table exactly as
 * b200surv_cox_fwd.  Integer sums: every rank obtains the bit-identical loss for any sharding.  epoch: 1, 2, 3, ...
 * the same sequence on every rank (slots are double-buffered on its parity); n, log_hz, time, event are the
 * rank-local rows.  Then b200surv_cox_bwd(mode BINNED) on the local rows. */
size_t b200surv_cox_peer_buffer_bytes(int32_t nbins);
/* Diagnostics: with B200SURV_PEER_TRACE set in the environment, b200surv_cox_binned_fwd_peer leaves 11 int64
 * %globaltimer stamps (ns; kernel start, pass 1, grid barrier, reduce, -, -, peers' sums added, look-back A, terms,
 * look-back G, header) at this byte offset of the workspace. */
size_t b200surv_cox_peer_trace_offset(int64_t n, int32_t nbins);
/* Peer-buffer plumbing for callers without their own symmetric allocator: alloc (cudaMalloc on the current
 * device, zero-filled) returns the pointer and a 64-byte CUDA IPC handle to hand to the other ranks of the
 * box (any byte transport); open maps a peer's handle into this process (peer access enabled lazily). */
#define B200SURV_PEER_HANDLE_BYTES 64
int32_t b200surv_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle);
int32_t b200surv_peer_open(const unsigned char *handle, void **dev_ptr);
int32_t b200surv_peer_close(void *dev_ptr);
int32_t b200surv_peer_free(void *dev_ptr);
int32_t b200surv_cox_binned_fwd_peer(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                                     int32_t ties, int32_t reduction, int32_t nbins, float shift,
                                     float *out_loss, void *state, size_t state_bytes, void *workspace,
                                     size_t workspace_bytes, void *const *peer_bufs, int32_t world,
                                     int32_t rank, uint32_t epoch, b200surv_stream_t stream);

/* SORTED mode over TIME-RANGE SHARDS for multi-GPU (path "Cox (B)"; BASELINE.json north_star: "an NCCL
 * allgather of per-shard (time, event, log_hz) boundary aggregates, with exact carry-in of each shard's scan prefix").
 * One cohort; shard r (rank r) holds n_r >= 1 rows in any order, every time of shard r <= every time of shard r + 1
 * (equal times may sit on both sides of an edge: tie groups that cross shards are merged exactly).  What the reference does
 * on one device with argsort + logcumsumexp (scripts/training/partial_modality_training.py:303-309) runs as four phases per
 * shard; between two phases the caller all-gathers ONE record of b200surv_cox_shard_record_bytes() (128) bytes per shard
 * (opaque; world * 128 bytes, shard order = time order) -- no other data crosses shards:
 *   _keys    sort keys, the shard's max log_hz / min / max time                                   -> rec0
 *   _sort    radix sort of the shard's rows (independent of rec0: overlap the all-gather with it)
 *   _reduce  (all rec0) common exponent shift = global max log_hz, edge times of the neighbours, order check; weights,
 *            per-tile sums, first tile scan; the shard's tile sequence folded into one element per chain   -> rec1
 *   _terms   (all rec1) carry-in: risk-set weight of the later shards, open tie groups at both edges;
 *            per-row Efron / Breslow terms, second tile scan                                              -> rec2
 *   _finish  (all rec2) carry-in of the prefix sums; loss, scale and header (identical on every shard), the per-row
 *            gradient of the shard's rows into `state`
 * then b200surv_cox_bwd(mode SORTED, n = n_r, n_seg = 1) on the shard's state.  workspace / state sizes:
 * b200surv_cox_workspace_bytes(n_r, 1, SORTED, 0) / b200surv_cox_state_bytes(n_r, 1, SORTED, 0).  The workspace carries
 * the shard's intermediate results from phase to phase.  world <= 64.  The result equals b200surv_cox_fwd(SORTED) on the
 * concatenated cohort up to fp64 summation order (fp32 loss / gradient: 1e-6 relative). */
size_t b200surv_cox_shard_record_bytes(void);
int32_t b200surv_cox_sorted_shard_keys(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                                       void *rec0_out, void *workspace, size_t workspace_bytes,
                                       b200surv_stream_t stream);
int32_t b200surv_cox_sorted_shard_sort(int64_t n, void *workspace, size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_cox_sorted_shard_reduce(const float *log_hz, int64_t n, const void *all_rec0, int32_t rank, int32_t world,
                                         void *rec1_out, void *workspace, size_t workspace_bytes,
                                         b200surv_stream_t stream);
int32_t b200surv_cox_sorted_shard_terms(int64_t n, int32_t ties, const void *all_rec1, int32_t rank, int32_t world,
                                        void *rec2_out, void *workspace, size_t workspace_bytes,
                                        b200surv_stream_t stream);
int32_t b200surv_cox_sorted_shard_finish(int64_t n, int32_t ties, int32_t reduction, const void *all_rec2, int32_t rank,
                                         int32_t world, float *out_loss, void *state, size_t state_bytes,
                                         void *workspace, size_t workspace_bytes, b200surv_stream_t stream);

/* Row-block shards -> time-range shards (the sample sort in front of the shard phases above): rank-local part.
 * b200surv_route_rows: dest[i] = number of `splitters` (device, n_dest - 1 ascending floats, the same on every rank) that are
 * <= time[i]; rows are grouped by destination, stable inside a destination:
 *   out_perm[j]     source row of the j-th routed row (int32; the way back for the gradient)
 *   out_counts[d]   rows for destination d (device int64[n_dest]; the caller reads them to size the all-to-all)
 *   out_log_hz / out_time / out_event [n]   the packed send buffers (any of them may be NULL)
 * The caller exchanges the groups (all-to-all with the counts as split sizes) and runs the shard phases on what it received.
 * Per step only log_hz changes: b200surv_route_gather(src, perm, n, out) packs a new vector with a kept permutation
 * (out[j] = src[perm[j]]), b200surv_route_scatter puts the returned gradient back (out[perm[j]] = src[j]).
 * workspace: b200surv_route_workspace_bytes(n).  n_dest <= 64.  Replaces, across GPUs, the reference's single-device
 * argsort(time) (scripts/training/partial_modality_training.py:303). */
size_t b200surv_route_workspace_bytes(int64_t n);
int32_t b200surv_route_rows(const float *log_hz, const float *time, const uint8_t *event, int64_t n,
                            const float *splitters, int32_t n_dest, float *out_log_hz, float *out_time,
                            uint8_t *out_event, int32_t *out_perm, int64_t *out_counts, void *workspace,
                            size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_route_gather(const float *src, const int32_t *perm, int64_t n, float *out, b200surv_stream_t stream);
int32_t b200surv_route_scatter(const float *src, const int32_t *perm, int64_t n, float *out, b200surv_stream_t stream);

/* ---- Harrell's concordance index: integer pair counts -------------------------------------- */
/* out_counts: int64[n_seg][6], ADDED to (caller zeroes): over rows i in [row_begin, row_end) of
 * each segment and all columns j of the same segment,
 *   strict pairs    (event_i && t_i <  t_j)            : [0] conc  [1] disc  [2] tied_risk
 *   same-time pairs (event_i && !event_j && t_i == t_j) : [3] conc  [4] disc  [5] tied_risk
 * with tie <=> fabsf(est_i - est_j) <= tied_tol (fp32), conc <=> !tie && est_j < est_i.
 * Row-block sharding across GPUs = disjoint [row_begin,row_end) per rank + an int64 SUM
 * all-reduce of the 6 counters (bit-exact, order independent).
 * algo 0 = direct all-pairs tiles (no preprocessing); algo 1 = sort by (time, event) first so that
 * every event row's comparable set is a suffix, then count pair by pair over upper-triangular tiles only;
 * algo 2 = algo 1's preprocessing, then RANKS instead of pairs: every 1024-column tile is also kept sorted by
 * estimate and the number of e_j below / up to a row's tie thresholds in a strictly-later tile is a binary search
 * (the same six integers bit for bit; 25x faster at 1M patients). */
size_t b200surv_cindex_workspace_bytes(int64_t n, int64_t n_seg, int32_t algo);
int32_t b200surv_cindex_counts(const float *estimate, const float *time, const uint8_t *event,
                               const int64_t *seg_offsets, int64_t n, int64_t n_seg,
                               int64_t row_begin, int64_t row_end, float tied_tol, int32_t algo,
                               int64_t *out_counts, void *workspace, size_t workspace_bytes,
                               b200surv_stream_t stream);
/* Strong scaling of ONE cohort over the GPUs of a box (every rank holds the full vectors): the event rows, in
 * (time, events first) order, are cut into tiles of 2048 rows that are dealt out round-robin; shard `shard` of
 * `n_shards` counts the pairs of its tiles against all columns.  The shards partition the pairs, so an int64 SUM
 * all-reduce of the six counters gives the single-GPU result bit for bit.  Unlike contiguous [row_begin,row_end)
 * blocks of the caller's row order, whose rows are scattered over the sorted order, a shard's tiles keep the
 * upper-triangular structure of the single-GPU run (the same fast-path share), and the triangle is balanced.
 * algo 1; out_counts is ADDED to.  b200surv_cindex_counts_shard_algo: the same with algo 1 or 2 (workspace of
 * b200surv_cindex_workspace_bytes(n, 1, algo) bytes). */
int32_t b200surv_cindex_counts_shard(const float *estimate, const float *time, const uint8_t *event, int64_t n,
                                     int32_t shard, int32_t n_shards, float tied_tol, int64_t *out_counts,
                                     void *workspace, size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_cindex_counts_shard_algo(const float *estimate, const float *time, const uint8_t *event, int64_t n,
                                          int32_t shard, int32_t n_shards, float tied_tol, int32_t algo,
                                          int64_t *out_counts, void *workspace, size_t workspace_bytes,
                                          b200surv_stream_t stream);
/* Many independent cohorts packed back to back (the CV sweep evaluates one C-index per fold and replica,
 * partial_modality_training.py:438-485 called per fold): cohort c is rows
 * [cohort_offsets_host[c], cohort_offsets_host[c+1]) -- a HOST array of n_cohorts+1 offsets -- and
 * out_counts is device int64[n_cohorts][6] (ADDED to).  One chain of launches per cohort; a workspace of k x
 * b200surv_cindex_workspace_bytes(n_max, 1, algo) bytes (rounded up to 256; k <= 8) lets k chains run at a time on
 * internal streams forked from and joined to `stream` with events (k = 1: one after the other on `stream`). */
int32_t b200surv_cindex_counts_cohorts(const float *estimate, const float *time, const uint8_t *event,
                                       const int64_t *cohort_offsets_host, int64_t n_cohorts,
                                       float tied_tol, int32_t algo, int64_t *out_counts, void *workspace,
                                       size_t workspace_bytes, b200surv_stream_t stream);

/* ---- fusion head: tensor-core GEMM primitive ----------------------------------------------- */
/* C[M][N] (fp32 and/or bf16 copy) = A * B (+ bias[n]) (ReLU), bf16 operands, fp32 accumulation in
 * TMEM (tcgen05.mma), operands staged by TMA.  Replaces the nn.Linear GEMMs of the heads
 * (partial_modality_training.py:196-232) and their gradients:
 *   a_mn == 0: A is row-major [M][K] (lda >= K)     a_mn == 1: A is row-major [K][M] (lda >= M)
 *   b_mn == 0: B is row-major [N][K] (ldb >= K)     b_mn == 1: B is row-major [K][N] (ldb >= N)
 * forward   y = x W^T + b : A = x [B][in], B = W [out][in]              (a_mn = 0, b_mn = 0)
 * dgrad     dx = dy W     : A = dy [B][out], B = W [out][in] as [K][N]  (a_mn = 0, b_mn = 1)
 * wgrad     dW = dy^T x   : A = dy [B][out] as [K][M], B = x [B][in] as [K][N]  (a_mn = 1, b_mn = 1)
 * Operand pointers must be 16-byte aligned with row pitches that are multiples of 8 elements; ragged
 * M, N, K are handled (TMA zero-fills out-of-bounds reads). */
int32_t b200surv_gemm_bf16(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb,
                           int32_t b_mn, int32_t M, int32_t N, int32_t K, float *c, int64_t ldc,
                           void *c_bf16, int64_t ldc_bf16, const float *bias, int32_t relu,
                           b200surv_stream_t stream);

/* The same GEMM with an explicit tile width and K split.  tile_n: 128 (what b200surv_gemm_bf16 uses), 192 or 256 columns
 * per 128-row output tile: wide tiles move fewer bytes from L2 per flop (rna_encoder.0 forward is bound by exactly that
 * at 128), 192 puts the rna_encoder.0 weight gradient (512 x 5005) on 108 CTAs = one wave of the 148 SMs; tile_n = 512
 * selects CTA PAIRS on 256 x 256 tiles (tcgen05.mma.cta_group::2, 2-CTA clusters): twice the math per byte a CTA loads.
 * splits >= 2:
 * K is cut into `splits` slices whose fp32 results [M][ldc] are written back to back into `slices` (c is ignored, no
 * bias / ReLU / bf16 copy); the caller sums them in slice order.  splits <= 1: like b200surv_gemm_bf16. */
int32_t b200surv_gemm_bf16_ex(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb, int32_t b_mn, int32_t M,
                              int32_t N, int32_t K, float *c, int64_t ldc, void *c_bf16, int64_t ldc_bf16, const float *bias,
                              int32_t relu, int32_t tile_n, int32_t splits, float *slices, b200surv_stream_t stream);

/* Diagnostics: CTA (0,0,0) of the CTA-pair GEMM writes %globaltimer stamps (start, first MMA, last commit, epilogue begin /
 * end, exit; from index 8 the arrival of the first 24 k-blocks) into buf (>= 32 int64 of device memory); NULL = off. */
void b200surv_debug_gemm_trace(long long *buf);

/* Split-K variant for outputs with few tiles and a long K (weight gradients): writes
 * b200surv_gemm_splitk_slices(M, N, K) (1..32) fp32 slices [M][ldc] back to back into `slices`; the caller sums them
 * in slice order (deterministic).  Same operand conventions as b200surv_gemm_bf16. */
int32_t b200surv_gemm_splitk_slices(int32_t M, int32_t N, int32_t K);
int32_t b200surv_gemm_bf16_splitk(const void *a, int64_t lda, int32_t a_mn, const void *b, int64_t ldb, int32_t b_mn,
                                  int32_t M, int32_t N, int32_t K, float *slices, int64_t ldc, b200surv_stream_t stream);

/* ---- fusion head: forward / backward -------------------------------------------------------- */
/* Parameters of the head, fp32 row-major [out][in], one pointer per reference state_dict entry
 * (PartialModalityNet: partial_modality_training.py:193-232; MultiModalSurvivalNet:
 * final_multimodal.py:95-120 has the same entries minus gate.*, which are then NULL):
 *   rna0 = rna_encoder.0 (512 x rna_dim), bn1 = rna_encoder.1 (512; running stats updated in training),
 *   rna4 = rna_encoder.4 (128 x 512), clin = clinical_encoder.0 (32 x 1), gate0 = gate.0 (64 x 291),
 *   gate2 = gate.2 (3 x 64), fus0 = fusion.0 (256 x 288), bn2 = fusion.1 (256), fus4 = fusion.4
 *   (128 x 256), cox = cox_head (1 x 128). */
typedef struct {
    const float *rna0_w, *rna0_b, *bn1_w, *bn1_b;
    float *bn1_rm, *bn1_rv;
    const float *rna4_w, *rna4_b, *clin_w, *clin_b, *gate0_w, *gate0_b, *gate2_w, *gate2_b, *fus0_w, *fus0_b,
        *bn2_w, *bn2_b;
    float *bn2_rm, *bn2_rv;
    const float *fus4_w, *fus4_b, *cox_w, *cox_b;
} b200surv_head_params;
/* Gradient outputs, same shapes as the parameters (written, not accumulated). */
typedef struct {
    float *rna0_w, *rna0_b, *bn1_w, *bn1_b, *rna4_w, *rna4_b, *clin_w, *clin_b, *gate0_w, *gate0_b, *gate2_w,
        *gate2_b, *fus0_w, *fus0_b, *bn2_w, *bn2_b, *fus4_w, *fus4_b, *cox_w, *cox_b;
} b200surv_head_grads;

size_t b200surv_head_saved_bytes(int64_t B, int32_t rna_dim);     /* activations kept fwd -> bwd */
size_t b200surv_head_workspace_bytes(int64_t B, int32_t rna_dim); /* scratch of one call         */

/* Forward.  ct_feat [B][128] (output of the CT encoder, contract a4), rna [B][rna_dim], clinical [B][1],
 * mask [B][3] = [image, rnaseq, clinical] or NULL for the ungated head (then gate must be NULL).
 * training != 0: BatchNorm uses batch statistics (B >= 2) and updates the running statistics, dropout
 * with probability dropout_p from a counter-based generator keyed by (seed, layer, element) -- backward
 * re-derives the same mask; keep1 [B][512] / keep2 [B][256] (nullable) export the keep masks.
 * training == B200SURV_HEAD_TRAIN_SEED_DEV: as training, but `seed` carries the ADDRESS of a uint64 in device memory
 * that the kernels read when they run -- a captured CUDA graph of fwd+bwd then draws a fresh dropout mask on every
 * replay once the caller bumps that word between replays.
 * Outputs: hazard [B], gate [B][3]. */
#define B200SURV_HEAD_TRAIN_SEED_DEV 2
/* OR-ed into `training`: the bf16 copy of `rna` already sits in the saved buffer (b200surv_head_stage_rna wrote it), `rna`
 * is not read.  A captured CUDA graph of the step then needs no copy of the 82 MB batch into a static input buffer: the
 * caller converts every new batch straight into the graph's saved buffer before the replay. */
#define B200SURV_HEAD_X_STAGED 4
/* OR-ed into `training` together with B200SURV_HEAD_TRAIN_SEED_DEV (forward only): the forward pass adds 1 to the
 * device-resident seed before anything reads it, so a replayed graph draws new dropout masks without a separate launch. */
#define B200SURV_HEAD_SEED_ADVANCE 8
int32_t b200surv_head_stage_rna(const float *rna, int64_t B, int32_t rna_dim, void *saved, size_t saved_bytes,
                                b200surv_stream_t stream);
/* The same plus up to three plain fp32 copies (copy_dst[k][0..copy_elems[k]) = copy_src[k][...]; HOST arrays of device
 * pointers) in the same launch: a new batch -- RNA matrix, CT features, age, modality mask -- reaches a captured step's
 * static buffers with one kernel instead of four (the batch-4 loop of partial_modality_training.py:382-400 hands over
 * exactly these four tensors per step). */
int32_t b200surv_head_stage_batch(const float *rna, int64_t B, int32_t rna_dim, void *saved, size_t saved_bytes,
                                  const float *const *copy_src, float *const *copy_dst, const int64_t *copy_elems,
                                  int32_t n_copies, b200surv_stream_t stream);
int32_t b200surv_head_fwd(const b200surv_head_params *params, const float *ct_feat, const float *rna,
                          const float *clinical, const float *mask, int64_t B, int32_t rna_dim,
                          int32_t training, float dropout_p, uint64_t seed, float *hazard, float *gate,
                          uint8_t *keep1, uint8_t *keep2, void *saved, size_t saved_bytes,
                          void *workspace, size_t workspace_bytes, b200surv_stream_t stream);

/* Backward of the same call (same B, rna_dim, training, dropout_p, seed, mask, clinical, saved).
 * d_hazard [B]; d_gate [B][3] or NULL (gradient flowing into the gate weights from outside, e.g. the
 * gate-entropy regulariser, partial_modality_training.py:322-331); outputs: grads, d_ct_feat [B][128]
 * (nullable). */
int32_t b200surv_head_bwd(const b200surv_head_params *params, const b200surv_head_grads *grads,
                          const float *d_hazard, const float *d_gate, const float *clinical,
                          const float *mask, int64_t B, int32_t rna_dim, int32_t training,
                          float dropout_p, uint64_t seed, float *d_ct_feat, const void *saved,
                          size_t saved_bytes, void *workspace, size_t workspace_bytes,
                          b200surv_stream_t stream);

/* Labelled-row compaction (partial_modality_training.py:401-408: hazard[mask], label[mask, 0], label[mask, 1] with
 * mask = has_survival): keeps the rows with has_survival[b] != 0 in order.  label is [B][2] = (time, event as 0/1
 * float).  Outputs sized for B rows: out_hazard, out_time, out_event (0/1 bytes = torch.bool storage), out_index
 * (source row of every kept row, for the backward scatter), out_counts[2] = {rows kept, events kept} (device; the
 * caller reads them to size the loss call and to apply the reference's skip rule n >= 2 && events > 0).
 * b200surv_scatter_rows: out_grad[B] = 0, out_grad[index[q]] = grad_sel[q] (the backward of the selection). */
size_t b200surv_compact_workspace_bytes(int64_t B);
int32_t b200surv_compact_labelled(const float *hazard, const float *label, const uint8_t *has_survival, int64_t B,
                                  float *out_hazard, float *out_time, uint8_t *out_event, int32_t *out_index,
                                  int64_t *out_counts, void *workspace, size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_scatter_rows(const float *grad_sel, const int32_t *index, int64_t n_sel, int64_t B, float *out_grad,
                              b200surv_stream_t stream);

/* Gate-entropy regulariser of the gated head (partial_modality_training.py:322-331, weighted 0.01 in the training
 * step :418-422): out_loss[0] = mean_b sum_k g log(g + eps) over gate [B][3]; backward d_gate [B][3] =
 * grad_out[0] * (log(g + eps) + g / (g + eps)) / B.  One launch each instead of eight small framework kernels. */
int32_t b200surv_gate_entropy_fwd(const float *gate, int64_t B, float eps, float *out_loss, b200surv_stream_t stream);
int32_t b200surv_gate_entropy_bwd(const float *gate, const float *grad_out, int64_t B, float eps, float *d_gate,
                                  b200surv_stream_t stream);

/* Fused clip_grad_norm_(max_norm) + Adam / AdamW step over a list of fp32 tensors -- the end of the reference's
 * training step (partial_modality_training.py:427-428 with optim.Adam(lr, weight_decay=1e-4) :536; simple_fusion.py
 * :273-274 with optim.AdamW :391).  params / grads / exp_avg / exp_avg_sq: HOST arrays of n_tensors device pointers,
 * numel_host their element counts; max_norm <= 0 disables clipping; adamw != 0: decoupled weight decay; step = 1, 2,
 * ... (bias correction); out_total_norm (device, nullable) receives the gradient norm before clipping.  Same update
 * as torch.optim.Adam / AdamW (no amsgrad), deterministic. */
size_t b200surv_clip_adam_workspace_bytes(const int64_t *numel_host, int32_t n_tensors);
int32_t b200surv_clip_adam_step(float *const *params, const float *const *grads, float *const *exp_avg,
                                float *const *exp_avg_sq, const int64_t *numel_host, int32_t n_tensors, float max_norm,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int32_t adamw,
                                int64_t step, float *out_total_norm, void *workspace, size_t workspace_bytes,
                                b200surv_stream_t stream);

/* ---- CT encoder feeding the head --------------------------------------------------------- */
/* The reference's non-MONAI CT branch (partial_modality_training.py:179-190): three Conv3d(k 3, stride 2, pad 1) +
 * BatchNorm3d + ReLU stages (1 -> 32 -> 64 -> 128 channels) and AdaptiveAvgPool3d(1), as primitives the host layer
 * (ctenc.py) strings together.  Activations are CHANNELS-LAST: rows = (sample, z, y, x) of a stage's output grid,
 * columns = channels; an input grid D x H x W gives an output grid ((D-1)/2+1) x ((H-1)/2+1) x ((W-1)/2+1).
 *   conv_first_fwd  : x fp32 [B][D][H][W] (one input channel), w fp32 [Cout][27] (torch (Cout,1,3,3,3)), bias ->
 *                     h fp32 [B*Do*Ho*Wo][Cout]                                       (direct kernel, K = 27)
 *   conv_first_wgrad: dw fp32 [Cout][27] = sum_rows dx[row][c] * patch(row)[tap]        (dx bf16 [rows][Cout])
 *   im2col / col2im : a bf16 [Bc*D*H*W][C] -> col bf16 [Bc*Do*Ho*Wo][27*C] (column = tap*C + c; the conv is then
 *                     b200surv_gemm_bf16(col, W_packed)); dcol bf16 -> da fp32 [Bc*D*H*W][C] by gathering (no atomics)
 *   weight_pack     : w fp32 (Cout,Cin,3,3,3) -> bf16 [Cout][27*Cin] tap-major; weight_unpack: the sum of `slices`
 *                     fp32 [Cout][27*Cin] gradient slices -> dw fp32 (Cout,Cin,3,3,3), added to dw when accumulate != 0
 *   bn_stats        : training: mu, rstd of the columns of x [R][C] (biased variance, eps 1e-5) and running <- 0.9
 *                     running + 0.1 (mu, unbiased variance) like nn.BatchNorm3d; eval: mu, rstd from the running stats
 *   bn_relu         : y bf16 = relu((x - mu) rstd gamma + beta);   bn_relu_pool: feat fp32 [B][C] = mean over the V
 *                     voxels of a sample of the same (stage 3 + AdaptiveAvgPool3d(1));  pool_bwd: dA[r][c] = dfeat[b][c]/V
 *   bn_bwd          : dy = [bn(x) > 0] dA; dgamma = sum dy xhat, dbeta = sum dy, dx bf16 = gamma rstd (dy - dbeta/R -
 *                     xhat dgamma/R) (training) or gamma rstd dy (eval); dbias = sum_rows dx (gradient of the conv bias)
 * C must be a power of two in [8, 256] for the bn_* calls and a multiple of 8 for im2col / col2im.  Every reduction
 * sums in a fixed order (deterministic).  workspace: b200surv_ct_workspace_bytes() bytes. */
size_t b200surv_ct_workspace_bytes(void);
int32_t b200surv_ct_conv_first_fwd(const float *x, const float *w, const float *bias, int64_t B, int32_t D, int32_t H,
                                   int32_t W, int32_t Cout, float *h, b200surv_stream_t stream);
int32_t b200surv_ct_conv_first_wgrad(const float *x, const void *dx_bf16, int64_t B, int32_t D, int32_t H, int32_t W,
                                     int32_t Cout, float *dw, void *workspace, size_t workspace_bytes,
                                     b200surv_stream_t stream);
int32_t b200surv_ct_im2col(const void *a_bf16, int64_t Bc, int32_t D, int32_t H, int32_t W, int32_t C, void *col_bf16,
                           b200surv_stream_t stream);
int32_t b200surv_ct_col2im(const void *dcol_bf16, int64_t Bc, int32_t D, int32_t H, int32_t W, int32_t C, float *da,
                           b200surv_stream_t stream);
int32_t b200surv_ct_weight_pack(const float *w, int32_t Cout, int32_t Cin, void *wr_bf16, b200surv_stream_t stream);
int32_t b200surv_ct_weight_unpack(const float *dwr_slices, int32_t slices, int32_t Cout, int32_t Cin, int32_t accumulate,
                                  float *dw, b200surv_stream_t stream);
int32_t b200surv_ct_bn_stats(const float *x, int64_t R, int32_t C, int32_t training, float *run_mean, float *run_var,
                             float *mu, float *rstd, void *workspace, size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_ct_bn_relu(const float *x, const float *mu, const float *rstd, const float *gamma, const float *beta,
                            int64_t R, int32_t C, void *y_bf16, b200surv_stream_t stream);
int32_t b200surv_ct_bn_relu_pool(const float *x, const float *mu, const float *rstd, const float *gamma,
                                 const float *beta, int64_t B, int32_t V, int32_t C, float *feat,
                                 b200surv_stream_t stream);
int32_t b200surv_ct_pool_bwd(const float *dfeat, int64_t B, int32_t V, int32_t C, float *dA, b200surv_stream_t stream);
int32_t b200surv_ct_bn_bwd(const float *x, const float *dA, const float *mu, const float *rstd, const float *gamma,
                           const float *beta, int64_t R, int32_t C, int32_t training, void *dx_bf16, float *dgamma,
                           float *dbeta, float *dbias, void *workspace, size_t workspace_bytes,
                           b200surv_stream_t stream);

/* The whole encoder (fixed architecture 1 -> 32 -> 64 -> 128 channels) in one call each way: ct fp32 [B][D][H][W],
 * feat / d_feat fp32 [B][128].  Parameters as in the reference state_dict (w[s] (Cout,Cin,3,3,3), b[s], gamma[s] =
 * BatchNorm weight, beta[s] = BatchNorm bias, running statistics updated in training); grads: one buffer per
 * parameter, overwritten.  `saved` (b200surv_ct_encoder_saved_bytes) carries the forward's activations to the
 * backward; `workspace` (b200surv_ct_encoder_workspace_bytes) is scratch.  No allocation, no synchronisation. */
typedef struct {
    const float *w[3], *b[3], *gamma[3], *beta[3];
    float *run_mean[3], *run_var[3];
} b200surv_ct_params;
typedef struct {
    float *w[3], *b[3], *gamma[3], *beta[3];
} b200surv_ct_grads;
size_t b200surv_ct_encoder_saved_bytes(int64_t B, int32_t D, int32_t H, int32_t W);
size_t b200surv_ct_encoder_workspace_bytes(int64_t B, int32_t D, int32_t H, int32_t W);
int32_t b200surv_ct_encoder_fwd(const float *ct, const b200surv_ct_params *p, int64_t B, int32_t D, int32_t H, int32_t W,
                                int32_t training, float *feat, void *saved, size_t saved_bytes, void *workspace,
                                size_t workspace_bytes, b200surv_stream_t stream);
int32_t b200surv_ct_encoder_bwd(const float *ct, const b200surv_ct_params *p, const float *d_feat, int64_t B, int32_t D,
                                int32_t H, int32_t W, int32_t training, const b200surv_ct_grads *grads, const void *saved,
                                size_t saved_bytes, void *workspace, size_t workspace_bytes, b200surv_stream_t stream);

/* ---- test hooks: the hand-written sort / scan primitives behind the SORTED Cox path and the C-index ---------- */
/* stable LSD radix sort of (u32 key, u32 value) pairs, in place (keys_tmp / vals_tmp: ping-pong buffers);
 * inclusive scan of (a, 2a, i), i combined by iop (0 add, 1 min, 2 max), ascending or descending index order. */
size_t b200surv_debug_sortscan_temp_bytes(int64_t n);
int32_t b200surv_debug_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                                  void *temp, b200surv_stream_t stream);
int32_t b200surv_debug_scan(const double *a, const int64_t *i, int64_t n, int32_t iop, int32_t reverse, double *out_a,
                            double *out_b, int64_t *out_i, void *temp, b200surv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SURV_H_ */
