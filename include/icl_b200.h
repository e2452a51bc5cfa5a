/*
 * icl_b200.h — C ABI of the B200-native `wgdi -icl` chaining engine.
 *
 * The reference (SunPengChuan/wgdi 0.75) is pure Python and has no FFI; the seam this library
 * replaces is the per-chromosome-pair call
 *
 *     collinearity(options, points).run()            wgdi/collinearity.py:5-70
 *
 * made from mycollinearity.single_pool (wgdi/run_colliearity.py:158-170) inside
 * multiprocessing workers (wgdi/run_colliearity.py:123-136).  Instead of one Python call per
 * chromosome pair, ALL pairs of a comparison are uploaded once and processed by CUDA kernels
 * (sm_100a).  The binding a maintainer of the reference would add is the ctypes stub shown in
 * INTEGRATION.md; wgdi_b200/_lib.py is the same stub as shipped here.
 *
 * Call sequence (one icl_ctx per GPU; a ctx is not thread-safe):
 *
 *   icl_create -> icl_upload -> icl_dp -> rank -> icl_extract -> icl_free_blocks ... -> icl_destroy
 *
 * where `rank` is either icl_rank_device (the order is computed on the device and never leaves it) or the host:
 * numpy.argsort of every pair segment of the scores icl_dp returned, handed to icl_extract.
 *
 * The split into icl_dp / icl_extract exists because the reference orders block extraction with
 * pandas' sort_values(by=['score1'], ascending=False) (wgdi/collinearity.py:49,60), i.e. numpy's
 * UNSTABLE argsort(kind='quicksort') whose tie permutation depends on the numpy build and the CPU
 * (SURVEY.md hazard H1).  Bit-exact output therefore needs that very permutation between the two
 * device phases: from numpy itself on the host, or from icl_rank_device, which reproduces numpy's three
 * argsort flavours exactly (the caller checks which one its numpy runs; wgdi_b200/engine.py does).
 *
 * All functions return ICL_OK (0) or a negative error code; icl_last_error() gives the message.
 * No exceptions cross the boundary.  Host input arrays are caller-owned and only read during the
 * call; icl_blocks arrays are owned by the context (see icl_extract).
 */
#ifndef ICL_B200_H
#define ICL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICL_OK 0
#define ICL_ERR_ARG (-1)    /* bad argument (null pointer, negative size, pair id out of range, ...) */
#define ICL_ERR_CUDA (-2)   /* a CUDA runtime call failed; message holds cudaGetErrorString */
#define ICL_ERR_STATE (-3)  /* call sequence violated (e.g. icl_extract before icl_dp) */
#define ICL_ERR_LIMIT (-4)  /* input exceeds a documented limit (see icl_upload / icl_dp) */
#define ICL_ERR_NOMEM (-5)  /* host allocation failed */

typedef struct icl_ctx icl_ctx;

/* Parameters of the DP engine: the typed subset of the `[collinearity]` options that
 * wgdi/collinearity.py:6-28 parses.  gap_penalty is fixed at -1 (collinearity.py:8). */
typedef struct icl_params {
    int32_t mg1;             /* `mg` first value: window on loc1, exclusive (collinearity.py:24,76-79); 2..255 */
    int32_t mg2;             /* `mg` second value: window on loc2; 2..255 */
    int32_t g0;              /* grading[0], divisor in the p value (collinearity.py:152) */
    int32_t over_gap;        /* minimum path length of a block, 3 in the reference (collinearity.py:13,129) */
    double coverage_ratio;   /* collinearity.py:16,129 (strict >) */
} icl_params;

/* Blocks in the reference's emission order: chromosome pair ascending (the caller numbers pairs in
 * the sorted (chr1, chr2) string order of wgdi/run_colliearity.py:118), forward blocks before
 * anti-diagonal blocks (collinearity.py:49-68), acceptance order inside a direction.
 * EVERY accepted block is returned (length >= over_gap); the caller applies
 *   p = round(p_raw, 4); drop if p > pvalue      wgdi/collinearity.py:153,55-56
 *   drop if m < 5                                wgdi/run_colliearity.py:178
 * because those filters do not feed back into the extraction. */
typedef struct icl_blocks {
    int64_t n_blocks;
    int64_t n_members;
    int32_t *pair;        /* [n_blocks] chromosome-pair id as given to icl_upload */
    int32_t *dir;         /* [n_blocks] 0 = forward (score1/path1), 1 = anti-diagonal (score2/path2) */
    int64_t *score;       /* [n_blocks] score of the block's top point (full path score) */
    int32_t *m;           /* [n_blocks] number of gene pairs in the block (len(self.path)) */
    int32_t *L;           /* [n_blocks] full path length (self.over_length) */
    int32_t *N;           /* [n_blocks] points inside the block's bounding box (collinearity.py:149) */
    int64_t *N1;          /* [n_blocks] sum of `times` inside the box before the bump (collinearity.py:148) */
    int32_t *L1;          /* [n_blocks] loc1 extent of the box */
    int32_t *L2;          /* [n_blocks] loc2 extent of the box */
    double *p_raw;        /* [n_blocks] unrounded p value `a` of collinearity.py:152 (binary64, left to right) */
    int64_t *member_off;  /* [n_blocks + 1] offsets into loc1/loc2/member_idx */
    int32_t *loc1;        /* [n_members] gene-order coordinate on genome 1, ascending inside a block */
    int32_t *loc2;        /* [n_members] gene-order coordinate on genome 2 */
    int32_t *member_idx;  /* [n_members] index of the point in the arrays given to icl_upload */
} icl_blocks;

/* Create a context on CUDA device `device`.  Fails with ICL_ERR_CUDA when no device is usable:
 * there is no CPU fallback. */
int icl_create(int device, icl_ctx **ctx);
void icl_destroy(icl_ctx *ctx);

/* Message of the most recent error on `ctx` (or of the last failed icl_create when ctx is NULL). */
const char *icl_last_error(const icl_ctx *ctx);

/* Page-locked host memory for the score / order arrays that cross PCIe between the two phases. */
int icl_host_alloc(icl_ctx *ctx, int64_t bytes, void **ptr);
int icl_host_free(icl_ctx *ctx, void *ptr);

/*
 * Copy the graded points of one comparison to the device (replaces the per-pair DataFrames that
 * wgdi/run_colliearity.py:117-136 pickles to its workers).  Arrays are in any order (normally BLAST
 * file order); pair_id[i] in [0, n_pairs).  The device orders the points of every pair by
 * (loc1, loc2) with a stable radix sort, i.e. equal coordinates keep upload order like the stable
 * lexsort of wgdi/run_colliearity.py:164-166.
 * Limits: 0 <= loc < 2^26, grading <= 255, n_points < 2^30, fewer than 2^24 points per pair,
 * bits(n_pairs) + bits(max loc1) + bits(max loc2) <= 64, sum over pairs of (loc1 range + 1) < 2^30 (all checked
 * here: ICL_ERR_LIMIT); icl_dp additionally requires that no window of mg1 consecutive rows of a pair holds more
 * than 65535 points.  The call also fixes the layout of the comparison (points and loc1 range of every pair, counted
 * on the device while the arrays are validated), so icl_dp needs no host round trip for it.
 */
int icl_upload(icl_ctx *ctx, int64_t n_points, int32_t n_pairs,
               const int32_t *pair_id, const int32_t *loc1, const int32_t *loc2, const int32_t *grading);

/*
 * icl_upload with the arrays in the narrowest unsigned element type that holds them: *_bytes = 1, 2 or 4 per element
 * (uint8 / uint16 / uint32-or-int32).  Gene orders per chromosome and pair ids normally fit 16 bits and gradings 8, so a
 * point crosses PCIe in 7 bytes instead of 16 — the upload is a fifth of the end-to-end time of a Vitis-sized comparison.
 * The arrays are widened on the device; everything else is as in icl_upload (same limits, same errors).
 */
int icl_upload_typed(icl_ctx *ctx, int64_t n_points, int32_t n_pairs, const void *pair_id, int32_t pair_bytes,
                     const void *loc1, int32_t loc1_bytes, const void *loc2, int32_t loc2_bytes,
                     const void *grading, int32_t grading_bytes);

/*
 * Grading classes for parameter sweeps over `grading` (BASELINE config 5).  The reference assigns a hit's grading
 * from the class of its per-query rank (wgdi/run_colliearity.py:30-33,64-86: rank < multiple -> grading[0], the next
 * four -> grading[1], the rest up to repeat_number -> grading[2]), so two confs that differ only in `grading` grade
 * the same hits into the same classes.  With a table set (n_classes > 0) the `grading` array given to icl_upload is
 * read as class ids in [0, n_classes) and every later icl_dp / icl_dp_begin chains with values[class]; the table may
 * be changed between icl_dp calls on the same upload.  n_classes = 0 (the default) restores plain grading values.
 * values must lie in 0..255.
 */
#define ICL_MAX_GRADE_CLASSES 8
int icl_set_grading_table(icl_ctx *ctx, int32_t n_classes, const int32_t *values);

/*
 * Sort + chaining DP in both directions (wgdi/collinearity.py:72-115: forward :74-93, anti-diagonal
 * :96-115).  Writes, for every pair segment pair_off[p]..pair_off[p+1] (see icl_layout), the final
 * scores of the segment's points in REVERSED (loc1, loc2) order into caller buffers of n_points elements.
 * Widened to int64 that is exactly the array pandas' nargsort(ascending=False) hands to numpy.argsort
 * (it reverses the column first; collinearity.py:49,60 -> pandas/core/sorting.py nargsort).  They cross
 * PCIe as int32 (scores are < 2^31); the caller MUST widen each segment to int64 before sorting because
 * numpy's sort kernels, and with them the tie permutation, differ per dtype and the reference's column is
 * int64.  May be called repeatedly on the same upload (e.g. a parameter sweep).
 */
int icl_dp(icl_ctx *ctx, const icl_params *params, int32_t *key_fwd, int32_t *key_rev);

/*
 * Streaming form of icl_dp: same work, but the call returns as soon as the DP kernel is launched and the scores
 * arrive item by item.  An item is one (pair, direction): id = dir * n_pairs + pair.  When an item is finished the
 * kernel writes its segment of key_fwd / key_rev (same layout and meaning as in icl_dp) straight into the caller's
 * buffers and then raises done_flags[id], so the host can run numpy's argsort on finished segments while heavier
 * chromosome pairs are still being chained (the reference does the same sort after each pair's DP,
 * wgdi/collinearity.py:49,60; here it overlaps the DP of the other pairs).
 *   key_fwd, key_rev, done_flags: page-locked memory from icl_host_alloc (the device writes to it directly);
 *   done_flags has 2 * n_pairs slots, all 0 when the call returns except those of empty pairs; a reader that
 *   sees done_flags[id] != 0 may read that item's segment (forward items: key_fwd, ids >= n_pairs: key_rev).
 * When the shared-memory DP kernel cannot be used (oversized window) the call degrades to icl_dp and raises all
 * flags before returning.  icl_dp_end waits for the device and must precede icl_extract; icl_layout may be called
 * in between (the layout is final once icl_dp_begin returns).
 */
int icl_dp_begin(icl_ctx *ctx, const icl_params *params, int32_t *key_fwd, int32_t *key_rev, int32_t *done_flags);
int icl_dp_end(icl_ctx *ctx);

/* Layout produced by icl_dp: pair_off[n_pairs + 1] (points of pair p are sorted positions
 * pair_off[p] .. pair_off[p+1]) and perm[n_points] (upload index of each sorted position).
 * Either pointer may be NULL. */
int icl_layout(icl_ctx *ctx, int64_t *pair_off, int32_t *perm);

/*
 * Rank every (pair, direction) item on the device instead of on the host, then call icl_extract with both argsort
 * pointers NULL.  icl_dp may then be called with both key pointers NULL (no score export).  Modes:
 *
 * ICL_RANK_NUMPY_SCALAR — the reference's tie order on every numpy that sorts int64 with its scalar introsort
 *   (npysort/quicksort.cpp `aquicksort_`: CPUs without AVX2 / AVX-512 sort dispatch, or NPY_DISABLE_CPU_FEATURES —
 *   which is how the end-to-end fixtures of this repository were generated).  The kernel reproduces that algorithm
 *   exactly: median of three, Hoare partition, insertion sort up to 16 elements, heap sort at the depth limit.
 *   A caller may use it whenever numpy.argsort in its own process behaves that way (wgdi_b200/engine.py checks this
 *   against the running numpy before choosing it); the result is then bit-identical to the host rank.
 * ICL_RANK_NUMPY_AVX512 — the same for numpy builds that dispatch int64 argsort to x86-simd-sort's AVX-512 kernels (hosts
 *   with AVX-512, e.g. this pool's): bitonic key-index networks up to 256 elements, 8-sample pivot, 4 x 8-lane unrolled
 *   compress-store partition, libstdc++ std::sort when the iteration budget runs out — all reproduced exactly.  The one
 *   branch that is not (std::sort's heap sort at its depth limit) makes the call fail with ICL_ERR_STATE: rank that
 *   comparison on the host instead.
 * ICL_RANK_NUMPY_AVX2 — the same for numpy builds / hosts where int64 argsort goes to x86-simd-sort's AVX2 kernels (x86 CPUs
 *   without AVX-512): the same algorithm on 4-lane vectors — groups of 16 in the partition, pivot = 3rd smallest of 4
 *   samples — and, because AVX2 has no compress-store (a permute and a full-width store take its place), the keys >=
 *   pivot of every vector stored in reversed lane order.
 * ICL_RANK_STABLE — NON-REFERENCE: a stable sort, i.e. what pandas' nargsort(ascending=False) would return if numpy's
 *   argsort were stable (reverse, stable ascending argsort, reverse): identical to the reference wherever scores are
 *   distinct, but tied scores come in reversed (loc1, loc2) order, so blocks with tied top scores can come out in a
 *   different order (and, where they share points, differently).  For throughput runs that do not need a tie order.
 */
#define ICL_RANK_STABLE 0
#define ICL_RANK_NUMPY_SCALAR 1
#define ICL_RANK_NUMPY_AVX512 2
#define ICL_RANK_NUMPY_AVX2 3
int icl_rank_device(icl_ctx *ctx, int32_t mode);

/*
 * Block extraction + p values (wgdi/collinearity.py:42-70, 117-153).
 * argsort_fwd / argsort_rev: for every pair segment the UNMODIFIED result of
 *     numpy.argsort(key[pair_off[p]:pair_off[p+1]].astype(int64), kind='quicksort')
 * on the arrays icl_dp returned, narrowed to int32 (segment-local indices).  The library undoes pandas' two
 * reversals itself.  The sort must be numpy's: it is unstable and its tie permutation decides which blocks
 * are extracted first (SURVEY.md hazard H1).
 * min_members: only blocks with at least that many gene pairs are returned (1 = everything the inner seam
 * `collinearity.run()` returns, 5 = what the file writer keeps, run_colliearity.py:178); smaller accepted
 * blocks still take part in the `times` accounting of the p value (collinearity.py:150).
 * The arrays of `out` live in page-locked staging owned by `ctx`: valid until the next icl_extract or
 * icl_destroy on that context; icl_free_blocks only clears the view.
 * Both argsort pointers NULL: use the order icl_rank_device left on the device.
 */
int icl_extract(icl_ctx *ctx, const int32_t *argsort_fwd, const int32_t *argsort_rev, int32_t min_members,
                icl_blocks *out);

/*
 * icl_extract with a choice of what crosses PCIe on the way back.  ICL_FIELDS_ALL: every array of `icl_blocks`, like icl_extract.
 * ICL_FIELDS_WRITER: what the collinearity file is written from (wgdi/run_colliearity.py:143-155,174-199: score, p value,
 * gene pairs of every block) — pair, score, m, p_raw, member_off, loc1, loc2; the other pointers of `out` are NULL.  Same
 * computation either way (N, N1 and the box extents still feed p_raw); 32 instead of 60 bytes per block and 8 instead of
 * 12 per gene pair come back, which matters when the blocks of several GPUs of a host go down at the same moment.
 */
#define ICL_FIELDS_ALL 0
#define ICL_FIELDS_WRITER 1
int icl_extract_fields(icl_ctx *ctx, const int32_t *argsort_fwd, const int32_t *argsort_rev, int32_t min_members,
                       int32_t fields, icl_blocks *out);
void icl_free_blocks(icl_blocks *blocks);

/* Test / diagnostics: DP state of one direction (0 forward, 1 anti-diagonal) in sorted point order.
 * pred = sorted position of the last updating source (-1 for path roots), depth = path length,
 * used = (usedtimes > 0).  Any pointer may be NULL. */
int icl_debug_dp(icl_ctx *ctx, int32_t dir, int32_t *score, int32_t *pred, int32_t *depth, uint8_t *used);

/* Test / diagnostics: the argsort icl_rank_device left on the device, forward segments then anti-diagonal ones (2 n
 * values, segment-local indices into the REVERSED score arrays — the format icl_extract takes from the host). */
int icl_debug_order(icl_ctx *ctx, int32_t *argsort);

/* Probe / test: numpy's argsort of flavour `mode` (one of ICL_RANK_NUMPY_*) of arbitrary int32 key segments, computed by
 * the same kernels and launch classes as icl_rank_device: keys[seg_off[j] .. seg_off[j+1]) -> argsort_out (segment-local
 * indices).  Host arrays; independent of the upload.  wgdi_b200/engine.py probes every launch size class with it against
 * the running numpy before it lets the device rank. */
int icl_debug_rank_keys(icl_ctx *ctx, int32_t mode, int32_t n_segments, const int64_t *seg_off, const int32_t *keys,
                        int32_t *argsort_out);

/* Test / verification: the key arrays the last icl_rank_device (numpy flavours) sorted — the scores of every item in
 * REVERSED point order, forward items then anti-diagonal ones (2 n values: what icl_dp exports when given buffers). */
int icl_debug_keys(icl_ctx *ctx, int32_t *keys);

/* Device time (ms, CUDA events on the context's stream) of the phases of the most recent
 * icl_dp / icl_extract calls; see ICL_T_* for the slots.  Returns the number of slots written. */
#define ICL_T_SORT 0      /* key build + radix sort + layout */
#define ICL_T_SETUP 1     /* row index, mirrored copy */
#define ICL_T_EDGES 2     /* candidate (edge) enumeration */
#define ICL_T_CHAIN 3     /* chaining DP, both directions */
#define ICL_T_EXPORT 4    /* score export kernel */
#define ICL_T_EXTRACT 5   /* block extraction */
#define ICL_T_PVALUE 6    /* times / p-value pass */
#define ICL_T_GATHER 7    /* output compaction + D2H */
#define ICL_T_D2H 8       /* scores device -> host */
#define ICL_T_COUNT 9
int icl_timings(icl_ctx *ctx, double *ms, int32_t n_slots);

/* Counters of the most recent calls: [0] edges enumerated, [1] kernels of this library launched by icl_dp (counted at
 * every launch, CUB excluded), [2] the same for icl_extract, [3] total dense rows, [4] largest window span (positions),
 * [5] DP kernel used (1 = shared-memory ring, 0 = generic), [6] kernels launched by icl_rank_device, [7] CUB device-wide
 * calls (radix sorts, scans) of the three together.  Returns the number of slots written (at most 8). */
int icl_counters(icl_ctx *ctx, int64_t *out, int32_t n_slots);

/* Device arrays and page-locked buffers of a destroyed (or regrown) context are not handed back to the driver: a
 * process-wide cache keeps them for the next context — cudaMalloc / cudaFree of a comparison's gigabytes cost far more than
 * its step, and drivers, batch runners and sharded runs create a context per comparison.  Limit: ICL_B200_CACHE_MB
 * (default 16384, 0 = no cache); a failed allocation empties the cache and tries again.  A recycled device block is
 * zeroed.  out[0] requests served from the cache, [1] requests that went to the driver, [2] bytes held now.  Returns the
 * number of slots written (at most 3); icl_cache_trim gives everything held back to the driver. */
int icl_cache_stats(int64_t *out, int32_t n_slots);
void icl_cache_trim(void);

/* The CUDA stream (cudaStream_t) every kernel and copy of `ctx` is issued on, for callers that bracket calls with
 * their own CUDA events (bench.py).  NULL in the emulation. */
void *icl_stream(icl_ctx *ctx);

/* "cuda" for the real library.  (A CPU emulation of the kernel bodies exists for the test-suite
 * only and reports "emu"; wgdi_b200 refuses to load it.) */
const char *icl_backend(void);

#ifdef __cplusplus
}
#endif
#endif /* ICL_B200_H */
