/*
 * CPU oracle for the `wgdi -icl` hot path, C restatement (TEST INFRASTRUCTURE, not product code).
 *
 * Same algorithm as oracle/icl_oracle.py (which is pinned against the unmodified reference by
 * tests/make_golden.py -> tests/golden/), written in plain C so that parity checks and the CPU
 * baseline of bench.py finish in seconds at the BASELINE.json sizes.  Nothing under wgdi_b200/
 * links or loads this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg do.
 *
 * Reference lines restated (paths relative to the reference checkout):
 *   icl_oracle_chain    wgdi/collinearity.py:72-93 (forward) and :96-115 (anti-diagonal)
 *   icl_oracle_extract  wgdi/collinearity.py:42-70, 117-144 (max_path) and :146-153 (p value)
 * The descending-score permutation (`order`) is NOT computed here: the reference obtains it from
 * numpy's unstable argsort through pandas (collinearity.py:49,60), so the caller passes the result
 * of that very numpy call (oracle/icl_oracle.py: nargsort_desc).
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* first index in [lo, hi) with a[idx] > v  (a ascending) */
static int64_t upper_bound32(const int32_t *a, int64_t lo, int64_t hi, int32_t v)
{
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] > v) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/*
 * One direction of score_matrix for one chromosome pair.
 * Points are sorted by (loc1, loc2) ascending (run_colliearity.py:164-166).
 * direction +1: sources in (loc1, loc2) order, candidates row < loc1_k < row+mg1 (collinearity.py:74-93)
 * direction -1: sources in (loc1 desc, loc2 asc) order, candidates row-mg1 < loc1_k < row, nearest row
 *               first (collinearity.py:96-115)
 * Outputs: score (int32), pred (index of last updating source, -1 for roots), depth (= len(path)),
 *          used (usedtimes > 0).
 */
void icl_oracle_chain(int64_t n, const int32_t *loc1, const int32_t *loc2, const int32_t *grading,
                      int32_t mg1, int32_t mg2, int32_t direction,
                      int32_t *score, int32_t *pred, int32_t *depth, uint8_t *used)
{
    if (n <= 0) return;
    /* distinct rows */
    int64_t nrows = 0;
    int64_t *rptr = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    int32_t *rval = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int64_t *rowof = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        if (i == 0 || loc1[i] != loc1[i - 1]) { rval[nrows] = loc1[i]; rptr[nrows] = i; nrows++; }
        rowof[i] = nrows - 1;
        score[i] = grading[i];          /* get_matrix, collinearity.py:35-36 */
        pred[i] = -1;
        depth[i] = 1;                   /* path = [label], :37 */
        used[i] = 0;
    }
    rptr[nrows] = n;

    for (int64_t step = 0; step < nrows; step++) {
        int64_t srow = direction > 0 ? step : nrows - 1 - step;
        for (int64_t i = rptr[srow]; i < rptr[srow + 1]; i++) {
            int32_t row = loc1[i], col = loc2[i];
            int32_t gap = mg2, row_old = row;                 /* :81 / :103 */
            int stop = 0;
            for (int64_t rr = srow + direction; rr >= 0 && rr < nrows && !stop; rr += direction) {
                int32_t rk = rval[rr];
                int32_t dr = (rk - row) * direction;
                if (dr >= mg1) break;                          /* exclusive bound :78 / :100 */
                int64_t hi = rptr[rr + 1];
                for (int64_t k = upper_bound32(loc2, rptr[rr], hi, col); k < hi && loc2[k] < col + mg2; k++) {
                    int32_t dc = loc2[k] - col;
                    if (dc > gap && (direction > 0 ? rk > row_old : rk < row_old)) { stop = 1; break; } /* :83-84 */
                    int32_t s = grading[k] - (dr + dc);        /* :85, gap_penalty = -1 */
                    if (s > 0 && score[k] < score[i] + s) {    /* :87 strict < */
                        score[k] = score[i] + s;
                        pred[k] = (int32_t)i;
                        used[i] = 1; used[k] = 1;
                        depth[k] = depth[i] + 1;
                        if (dc < gap) gap = dc;                /* :92 */
                        row_old = rk;                          /* :93 */
                    }
                }
            }
        }
    }
    free(rptr); free(rval); free(rowof);
}

/*
 * Greedy extraction for one direction of one pair (collinearity.py:49-57 / :60-68 with max_path :117-144
 * and p_value_estimated :146-153).  `order` = descending-score permutation over all n points.
 * `times` (n, initialised to 1 by the caller, collinearity.py:34) is shared by both directions.
 * Every ACCEPTED block is written (the caller applies the rounded-p filter of :55-56 and the N>=5 filter
 * of run_colliearity.py:178, neither of which feeds back into the extraction).
 * Members are stored sorted by loc1 ascending (:140).  Returns the number of blocks.
 */
int64_t icl_oracle_extract(int64_t n, const int32_t *loc1, const int32_t *loc2,
                           const int32_t *score, const int32_t *pred, const int32_t *depth,
                           const uint8_t *used, const int32_t *order, int32_t *times,
                           int32_t g0, int32_t over_gap, double coverage_ratio,
                           int64_t *blk_off, int32_t *members,
                           int32_t *blk_top, int32_t *blk_L, int32_t *blk_N, int64_t *blk_N1,
                           int32_t *blk_L1, int32_t *blk_L2, double *blk_a)
{
    uint8_t *present = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    int32_t *rem = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    memcpy(present, used, (size_t)n);                           /* drop usedtimes < 1, :50/:61 */
    int64_t nb = 0, nm = 0;
    blk_off[0] = 0;
    for (int64_t t = 0; t < n; t++) {
        int32_t top = order[t];
        if (!present[top]) continue;
        int32_t L = depth[top];
        int64_t m = 0;
        for (int32_t q = top; q >= 0; q = pred[q])              /* path members still in the frame, :125 */
            if (present[q]) rem[m++] = q;
        if (L >= over_gap && (double)m / (double)L > coverage_ratio) {      /* :129 */
            int32_t l1min = loc1[rem[0]], l1max = l1min, l2min = loc2[rem[0]], l2max = l2min;
            for (int64_t j = 0; j < m; j++) {
                int32_t q = rem[j];
                present[q] = 0;                                 /* :130 */
                if (loc1[q] < l1min) l1min = loc1[q];
                if (loc1[q] > l1max) l1max = loc1[q];
                if (loc2[q] < l2min) l2min = loc2[q];
                if (loc2[q] > l2max) l2max = loc2[q];
            }
            int64_t N = 0, N1 = 0;
            int64_t j0 = 0, j1 = n;                             /* first point with loc1 >= l1min */
            while (j0 < j1) { int64_t mid = j0 + ((j1 - j0) >> 1); if (loc1[mid] >= l1min) j1 = mid; else j0 = mid + 1; }
            for (int64_t j = j0; j < n; j++) {                  /* :134-137, inclusive box over ALL points */
                if (loc1[j] > l1max) break;
                if (loc2[j] >= l2min && loc2[j] <= l2max) { N1 += times[j]; N++; times[j]++; }   /* :148-150 */
            }
            double L1 = (double)(l1max - l1min + 1), L2 = (double)(l2max - l2min + 1), dm = (double)m;
            /* :152, left to right in binary64 */
            volatile double a = 1.0 - (double)score[top] / dm / (double)g0;
            a = a * (double)(N1 - m + 1);
            a = a / (double)N;
            a = a * (L1 - dm + 1.0);
            a = a * (L2 - dm + 1.0);
            a = a / L1;
            a = a / L2;
            /* members by loc1 ascending: the walk top->root is loc1-descending forward and ascending on the
               anti-diagonal; loc1 is strictly monotone along a path, so a reversal check suffices */
            if (m > 1 && loc1[rem[0]] > loc1[rem[m - 1]])
                for (int64_t j = 0; j < m; j++) members[nm + j] = rem[m - 1 - j];
            else
                for (int64_t j = 0; j < m; j++) members[nm + j] = rem[j];
            nm += m;
            blk_top[nb] = top; blk_L[nb] = L; blk_N[nb] = (int32_t)N; blk_N1[nb] = N1;
            blk_L1[nb] = l1max - l1min + 1; blk_L2[nb] = l2max - l2min + 1; blk_a[nb] = a;
            nb++;
            blk_off[nb] = nm;
        } else {
            present[top] = 0;                                   /* :143 */
        }
    }
    free(present); free(rem);
    return nb;
}

/* ---------------------------------------------------------------------------------------------
 * Batch drivers: the reference distributes chromosome pairs over worker processes
 * (run_colliearity.py:123-136); here a pthread pool pulls (pair, direction) items from an atomic
 * counter.  Arrays are the concatenation of all pairs; off[p]..off[p+1] is pair p.
 * ------------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct {
    int32_t n_items;
    volatile int32_t next;
    void (*fn)(int32_t item, void *arg);
    void *arg;
} pool_t;

static void *pool_worker(void *vp)
{
    pool_t *p = (pool_t *)vp;
    for (;;) {
        int32_t it = __sync_fetch_and_add(&p->next, 1);
        if (it >= p->n_items) break;
        p->fn(it, p->arg);
    }
    return 0;
}

static void pool_run(int32_t n_items, int32_t n_threads, void (*fn)(int32_t, void *), void *arg)
{
    pool_t p = { n_items, 0, fn, arg };
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int32_t t = 1; t < n_threads; t++) pthread_create(&th[t], 0, pool_worker, &p);
    pool_worker(&p);
    for (int32_t t = 1; t < n_threads; t++) pthread_join(th[t], 0);
}

typedef struct {
    const int32_t *order;          /* items in descending size (longest first) or NULL */
    const int64_t *off; const int32_t *loc1, *loc2, *grading; int32_t mg1, mg2;
    int32_t *score[2], *pred[2], *depth[2]; uint8_t *used[2];
} chain_args;

static void chain_item(int32_t it, void *va)
{
    chain_args *c = (chain_args *)va;
    if (c->order) it = c->order[it];
    int32_t p = it >> 1, d = it & 1;
    int64_t a = c->off[p], n = c->off[p + 1] - a;
    icl_oracle_chain(n, c->loc1 + a, c->loc2 + a, c->grading + a, c->mg1, c->mg2, d ? -1 : +1,
                     c->score[d] + a, c->pred[d] + a, c->depth[d] + a, c->used[d] + a);
}

void icl_oracle_chain_batch(int32_t n_pairs, int32_t n_threads, const int32_t *item_order,
                            const int64_t *off, const int32_t *loc1, const int32_t *loc2,
                            const int32_t *grading, int32_t mg1, int32_t mg2,
                            int32_t *score_f, int32_t *pred_f, int32_t *depth_f, uint8_t *used_f,
                            int32_t *score_r, int32_t *pred_r, int32_t *depth_r, uint8_t *used_r)
{
    chain_args c = { item_order, off, loc1, loc2, grading, mg1, mg2,
                     { score_f, score_r }, { pred_f, pred_r }, { depth_f, depth_r }, { used_f, used_r } };
    pool_run(2 * n_pairs, n_threads, chain_item, &c);
}

/* Extraction for all pairs: both directions of a pair run on one thread because they share `times`
 * (collinearity.py:34,150).  order_f/order_r hold pair-local indices.  Layout of the outputs:
 *   members : 2*off[p] .. (forward area of n slots, then anti-diagonal area of n slots)
 *   block records : blk_base[p] .. (forward blocks, then anti-diagonal blocks)
 *   blk_off : one slot more than blocks per direction -> starts at blk_base[p] + 2p (fwd) and
 *             blk_base[p] + n_blk_f[p] + 2p + 1 (rev); values are relative to the direction's member area */
typedef struct {
    const int32_t *pair_order;
    const int64_t *off; const int32_t *loc1, *loc2;
    const int32_t *score[2], *pred[2], *depth[2]; const uint8_t *used[2]; const int32_t *order[2];
    int32_t g0, over_gap; double cov;
    const int64_t *blk_base; int64_t *n_blk[2];
    int64_t *blk_off; int32_t *members, *blk_top, *blk_L, *blk_N; int64_t *blk_N1; int32_t *blk_L1, *blk_L2; double *blk_a;
} extract_args;

static void extract_item(int32_t p, void *va)
{
    extract_args *c = (extract_args *)va;
    if (c->pair_order) p = c->pair_order[p];
    int64_t a = c->off[p], n = c->off[p + 1] - a;
    int32_t *times = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    for (int64_t j = 0; j < n; j++) times[j] = 1;              /* get_matrix, collinearity.py:34 */
    int64_t b = c->blk_base[p];
    for (int d = 0; d < 2; d++) {
        int64_t nb = icl_oracle_extract(n, c->loc1 + a, c->loc2 + a, c->score[d] + a, c->pred[d] + a, c->depth[d] + a,
                                        c->used[d] + a, c->order[d] + a, times, c->g0, c->over_gap, c->cov,
                                        c->blk_off + b + 2 * p + d, c->members + 2 * a + d * n,
                                        c->blk_top + b, c->blk_L + b, c->blk_N + b, c->blk_N1 + b,
                                        c->blk_L1 + b, c->blk_L2 + b, c->blk_a + b);
        c->n_blk[d][p] = nb;
        b += nb;
    }
    free(times);
}

void icl_oracle_extract_batch(int32_t n_pairs, int32_t n_threads, const int32_t *pair_order,
                              const int64_t *off, const int32_t *loc1, const int32_t *loc2,
                              const int32_t *score_f, const int32_t *pred_f, const int32_t *depth_f, const uint8_t *used_f,
                              const int32_t *score_r, const int32_t *pred_r, const int32_t *depth_r, const uint8_t *used_r,
                              const int32_t *order_f, const int32_t *order_r,
                              int32_t g0, int32_t over_gap, double coverage_ratio,
                              const int64_t *blk_base, int64_t *n_blk_f, int64_t *n_blk_r,
                              int64_t *blk_off, int32_t *members,
                              int32_t *blk_top, int32_t *blk_L, int32_t *blk_N, int64_t *blk_N1,
                              int32_t *blk_L1, int32_t *blk_L2, double *blk_a)
{
    extract_args c = { pair_order, off, loc1, loc2,
                       { score_f, score_r }, { pred_f, pred_r }, { depth_f, depth_r }, { used_f, used_r },
                       { order_f, order_r }, g0, over_gap, coverage_ratio, blk_base, { n_blk_f, n_blk_r },
                       blk_off, members, blk_top, blk_L, blk_N, blk_N1, blk_L1, blk_L2, blk_a };
    pool_run(n_pairs, n_threads, extract_item, &c);
}
