/* Harrell's concordance index -- CPU oracle in plain C.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py; parity with torchsurv is UNPINNED, the published pair rule is
 * restated and pinned against the reference's runnable fallback through tests/golden/).
 *
 * Boundary mirrored: ConcordanceIndex()(estimate, event, time) as called at
 *   scripts/training/partial_modality_training.py:290-294, simple_fusion.py:330-331.
 * Reference fallback it is pinned against: simple_fusion.py:59-73 (double loop:
 *   event[i] && time[j] > time[i] -> permissible; log_hazard[i] > log_hazard[j] -> concordant).
 *
 * Six int64 counters:
 *   strict pairs     (event_i && t_i <  t_j):             out[0]=conc out[1]=disc out[2]=tied_risk
 *   same-time pairs  (event_i && !event_j && t_i == t_j): out[3]=conc out[4]=disc out[5]=tied_risk
 * with, for an ordered pair (i, j):  tie  <=> fabsf(est_i - est_j) <= tied_tol   (fp32 arithmetic,
 * exactly what a float32 tensor expression |est - est_i| <= tied_tol evaluates), conc <=> !tie &&
 * est_j < est_i, disc otherwise.
 *   fallback C-index (simple_fusion.py:59-73) = conc_strict_notol / (all strict pairs)
 *   Harrell / scikit-survival / torchsurv-style = (C + T/2) / (C + D + T), summed over both groups.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int is_tie(float a, float b, float tol) {
    volatile float d = a - b; /* volatile: force a rounded fp32 difference, no excess precision */
    return fabsf(d) <= tol;
}

/* O(n^2) literal definition over rows [row_begin, row_end) x all columns. */
void cindex_counts_brute(const float *est, const float *time, const uint8_t *event, int64_t n,
                         float tied_tol, int64_t row_begin, int64_t row_end, int64_t out[6]) {
    int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : c0, c1, c2, c3, c4, c5)
    for (int64_t i = row_begin; i < row_end; ++i) {
        if (!event[i]) continue;
        const float ti = time[i], ei = est[i];
        for (int64_t j = 0; j < n; ++j) {
            const float tj = time[j];
            int strict = tj > ti;
            int same = (tj == ti) && !event[j];
            if (!(strict || same)) continue;
            int tie = is_tie(ei, est[j], tied_tol);
            int conc = !tie && (est[j] < ei);
            if (strict) {
                if (tie) c2++; else if (conc) c0++; else c1++;
            } else {
                if (tie) c5++; else if (conc) c3++; else c4++;
            }
        }
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3; out[4] = c4; out[5] = c5;
}

/* ---------- O(n log n): Fenwick tree over estimate ranks, sweep from the latest time ---------- */
typedef struct { float t; int64_t i; } tkey_t;
static int cmp_tkey(const void *a, const void *b) {
    const tkey_t *x = (const tkey_t *)a, *y = (const tkey_t *)b;
    if (x->t < y->t) return -1;
    if (x->t > y->t) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}
static void fen_add(int64_t *f, int64_t n, int64_t k) { for (k++; k <= n; k += k & -k) f[k]++; }
static int64_t fen_sum(const int64_t *f, int64_t k) { /* count of ranks < k */
    int64_t s = 0;
    for (; k > 0; k -= k & -k) s += f[k];
    return s;
}
/* first index in sorted u[0..n) with u[k] >= v */
static int64_t lower_bound(const float *u, int64_t n, float v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (u[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

int cindex_counts_fast(const float *est, const float *time, const uint8_t *event, int64_t n,
                       float tied_tol, int64_t out[6]) {
    memset(out, 0, 6 * sizeof(int64_t));
    if (n <= 1) return 0;
    tkey_t *ord = (tkey_t *)malloc(sizeof(tkey_t) * n);
    float *u = (float *)malloc(sizeof(float) * n);
    int64_t *fen = (int64_t *)calloc(n + 1, sizeof(int64_t));
    int64_t *lo_rank = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *hi_rank = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *my_rank = (int64_t *)malloc(sizeof(int64_t) * n);
    if (!ord || !u || !fen || !lo_rank || !hi_rank || !my_rank) return -1;
    for (int64_t i = 0; i < n; ++i) { ord[i].t = time[i]; ord[i].i = i; u[i] = est[i]; }
    qsort(ord, n, sizeof(tkey_t), cmp_tkey);
    qsort(u, n, sizeof(float), cmp_float);
    /* tie interval of row i in rank space: ranks [lo_rank, hi_rank) are the u[k] with
     * fabsf(est_i - u[k]) <= tol; the predicate is monotone on each side of est_i because fp32
     * subtraction is monotone, so both ends are found by bisection on the exact predicate. */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float e = est[i];
        int64_t p = lower_bound(u, n, e); /* u[p] == e */
        my_rank[i] = p;
        int64_t lo = 0, hi = p; /* smallest k in [0,p] with tie(u[k]); tie(u[p]) holds */
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (is_tie(e, u[mid], tied_tol)) hi = mid; else lo = mid + 1; }
        lo_rank[i] = lo;
        lo = p; hi = n; /* first k in [p, n] with !tie */
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (is_tie(e, u[mid], tied_tol)) lo = mid + 1; else hi = mid; }
        hi_rank[i] = lo;
    }
    /* sweep distinct times from the latest to the earliest; the tree holds every row with a
     * strictly later time when a group's event rows are queried (strict pairs), then the group's
     * censored rows are added and the same event rows queried again (difference = same-time pairs). */
    int64_t inserted = 0;
    int64_t g1 = n;
    while (g1 > 0) {
        int64_t g0 = g1 - 1;
        while (g0 > 0 && ord[g0 - 1].t == ord[g1 - 1].t) g0--;
        for (int64_t k = g0; k < g1; ++k) {
            int64_t i = ord[k].i;
            if (!event[i]) continue;
            int64_t below = fen_sum(fen, lo_rank[i]);
            int64_t upto = fen_sum(fen, hi_rank[i]);
            out[0] += below; out[2] += upto - below; out[1] += inserted - upto;
            /* the second query below counts strict + same-time; pre-subtract the strict part */
            out[3] -= below; out[5] -= upto - below; out[4] -= inserted - upto;
        }
        for (int64_t k = g0; k < g1; ++k) {
            int64_t i = ord[k].i;
            if (!event[i]) { fen_add(fen, n, my_rank[i]); inserted++; }
        }
        for (int64_t k = g0; k < g1; ++k) {
            int64_t i = ord[k].i;
            if (!event[i]) continue;
            int64_t below = fen_sum(fen, lo_rank[i]);
            int64_t upto = fen_sum(fen, hi_rank[i]);
            out[3] += below; out[5] += upto - below; out[4] += inserted - upto;
        }
        for (int64_t k = g0; k < g1; ++k) {
            int64_t i = ord[k].i;
            if (event[i]) { fen_add(fen, n, my_rank[i]); inserted++; }
        }
        g1 = g0;
    }
    free(ord); free(u); free(fen); free(lo_rank); free(hi_rank); free(my_rank);
    return 0;
}
