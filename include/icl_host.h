/*
 * icl_host.h — native host-side helpers of the `-icl` path (ingest and writer).
 *
 * Plain C ABI, no CUDA: libicl_host.so is built with g++ and is independent of libicl_b200.so.
 * What each entry point replaces in the reference:
 *
 *   icl_host_read_m8       wgdi/base.py:172-183 (newblast: pandas.read_csv + the row filters + drop_duplicates)
 *                          and the id -> gene lookups of wgdi/run_colliearity.py:103-106
 *   icl_host_grade         wgdi/run_colliearity.py:59-86 (deal_blast_for_genomes), except for queries whose tied
 *                          bit scores straddle a class boundary: those are reported and re-ranked by the caller
 *                          with numpy's own unstable argsort (SURVEY.md hazard H2)
 *   icl_host_ingest_*      the whole ingest of `comparison = genomes` in one handle: icl_host_read_m8 + icl_host_grade +
 *                          the point table of wgdi/run_colliearity.py:103-121,164-166 (chromosome pair, loc1, loc2,
 *                          grading per kept hit), so that nothing per hit is touched by Python
 *   icl_host_write_blocks  wgdi/run_colliearity.py:143-155,181-199 (gene names, strands, text assembly, file)
 *
 * All functions return 0 or a negative code and fill `err` (NUL-terminated, truncated to errlen).
 */
#ifndef ICL_HOST_H
#define ICL_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct icl_hits {
    int64_t n;             /* rows that survive the filters of wgdi/base.py:178-180 */
    int64_t n_rows_file;   /* data rows in the file (blank lines are skipped like pandas does) */
    int32_t bits_is_int;   /* 1 when every column-11 token is an integer literal: pandas infers int64, else float64 */
    int64_t *row;          /* [n] 0-based row number in the file = the reference's row label */
    int32_t *g1;           /* [n] row of id1 in the genome-1 gene table */
    int32_t *g2;           /* [n] row of id2 in the genome-2 gene table */
    double *bits;          /* [n] column 11 (bit score) */
} icl_hits;

/*
 * Parse a BLAST -outfmt 6 / m8 table (tab separated, >= 12 columns, no header) with `threads` threads and keep
 * the rows the reference keeps: bitscore >= min_score and evalue < max_evalue and id1 != id2 (base.py:178),
 * both ids present in the gene tables (:179), first occurrence of every (id1, id2) in file order (:180).
 * reverse != 0 swaps the two id columns first (base.py:176-177; the caller passes 1 only for the exact string
 * 'true').  Gene tables: concatenated id bytes + offsets[n+1].
 * Numbers: exact (correctly rounded) conversion; identical to pandas' parser for tokens of up to 15 significant
 * digits, which covers everything BLAST prints (pandas' default converter is itself an ulp off for some longer ones
 * and for very large exponents, e.g. it reads "1.00e-300" as 9.999999999999999e-301 — relevant only if the e-value
 * limit itself were such a number).
 */
int icl_host_read_m8(const char *path, int32_t reverse, double min_score, double max_evalue,
                     const char *ids1, const int64_t *off1, int64_t n1,
                     const char *ids2, const int64_t *off2, int64_t n2,
                     int32_t threads, icl_hits *out, char *err, int32_t errlen);
void icl_host_free_hits(icl_hits *hits);

/*
 * Grading by per-query hit rank (run_colliearity.py:59-86): for every query (all target chromosomes together)
 * hits are ranked by bit score descending; rank < multiple -> grading[0]; rank < multiple+4 -> grading[1];
 * rank < repeat_number -> grading[2]; others 0.  Ties are broken by file order here; a query whose class
 * assignment depends on the tie order (equal scores on both sides of a boundary) is flagged in
 * ambiguous[query] so that the caller can apply numpy's permutation to exactly those queries.
 */
int icl_host_grade(int64_t n, const int32_t *query, const double *bits, int32_t n_query,
                   int32_t multiple, int32_t repeat_number, const int32_t *grading3,
                   int32_t *grade_out, uint8_t *ambiguous, int32_t threads, char *err, int32_t errlen);

/*
 * Per-query hit rank for `-bi` (`group.sort_values(by=11, ascending=False)[:repeat_number]`, wgdi/block_info.py:104-105):
 * rank_out[i] = 0-based rank of hit i among its query's hits by bit score descending, -1 beyond the first repeat_number.
 * Ties are broken by file order here; ambiguous[query] = 1 when a run of tied scores touches a rank below class_bound (the
 * homo classes of :106-118 switch at ranks 1..10) or the cut at repeat_number: the caller re-ranks those queries with
 * numpy's own unstable argsort (SURVEY.md hazard H2).
 */
int icl_host_rank(int64_t n, const int32_t *query, const double *bits, int32_t n_query, int32_t repeat_number,
                  int32_t class_bound, int64_t *rank_out, uint8_t *ambiguous, int32_t threads, char *err, int32_t errlen);

/*
 * Fused ingest of one `comparison = genomes` run (wgdi/base.py:172-183 + wgdi/run_colliearity.py:59-86,103-121).
 *
 *   icl_host_ingest_open     parses and filters like icl_host_read_m8, removes the duplicates and grades like
 *                            icl_host_grade with classes 1 / 2 / 3 (blue_extra = 4 for -icl: bluenum = 4 + rednum, :64).
 *                            Queries whose classes depend on the order of tied bit scores are NOT graded: their hits
 *                            (duplicates removed, file order) are set aside for the caller.
 *   icl_host_ingest_ambiguous  amb_off[n_amb_queries + 1], amb_bits[n_amb_hits]: the bit scores of those queries, in the
 *                            order `group.sort_values(by=11)` of :67 sees them.
 *   icl_host_ingest_set_classes  amb_cls[n_amb_hits] in {0 (dropped), 1, 2, 3}: what the caller's numpy argsort — the
 *                            reference's own unstable sort, hazard H2 — makes of them.
 *   icl_host_ingest_count    present[nc1 * nc2] = 1 where a graded hit falls on (chromosome code 1, chromosome code 2);
 *                            n_points = graded hits (class > 0).  chr1 / chr2: chromosome code per gene row.
 *   icl_host_ingest_points   the points in file order: pair_id = pair_rank[chr1 * nc2 + chr2] (the caller numbers the
 *                            present pairs in the string order of :118), loc1 / loc2 = order of the two genes,
 *                            grading = grading3[class - 1], grade_class = class - 1.  Needs icl_host_ingest_count first.
 */
typedef struct icl_ingest icl_ingest;
typedef struct icl_ingest_info {
    int64_t n_hits;        /* rows that survive the filters of wgdi/base.py:178-180 */
    int64_t n_rows_file;
    int32_t bits_is_int;   /* as in icl_hits: the caller ranks the ambiguous queries on int64 keys when pandas would */
    int64_t n_amb_queries;
    int64_t n_amb_hits;
} icl_ingest_info;

int icl_host_ingest_open(const char *path, int32_t reverse, double min_score, double max_evalue,
                         const char *ids1, const int64_t *off1, int64_t n1,
                         const char *ids2, const int64_t *off2, int64_t n2,
                         int32_t multiple, int32_t repeat_number, int32_t blue_extra,
                         int32_t threads, icl_ingest **out, icl_ingest_info *info, char *err, int32_t errlen);
int icl_host_ingest_ambiguous(const icl_ingest *h, int64_t *amb_off, double *amb_bits);
int icl_host_ingest_set_classes(icl_ingest *h, const int8_t *amb_cls);
int icl_host_ingest_count(icl_ingest *h, const int32_t *chr1, const int32_t *chr2, int32_t nc1, int32_t nc2,
                          uint8_t *present, int64_t *n_points);
int icl_host_ingest_points(const icl_ingest *h, const int32_t *chr1, const int64_t *order1, const int32_t *chr2,
                           const int64_t *order2, int32_t nc2, const int32_t *pair_rank, const int32_t *grading3,
                           int32_t *pair_id, int32_t *loc1, int32_t *loc2, int32_t *grading, int32_t *grade_class);
void icl_host_ingest_close(icl_ingest *h);

/*
 * Write the collinearity file.  For every selected block b (already filtered and ordered by the caller):
 *   header[b]                                   "# Alignment k: score=.. pvalue=.. N=.. chr1&chr2 plus|minus"
 *   one line per member   name1 loc1 name2 loc2 strand      strand = "1" iff the two strand strings are equal
 * Genes are found by (chromosome code, order): row = row_of[code * maxo + order].
 */
int icl_host_write_blocks(const char *path, int64_t n_sel,
                          const char *headers, const int64_t *header_off,
                          const int64_t *mem_begin, const int64_t *mem_end,
                          const int32_t *loc1, const int32_t *loc2,
                          const int32_t *chr1_code, const int32_t *chr2_code,
                          const int64_t *row_of1, int64_t maxo1, const char *names1, const int64_t *name_off1,
                          const int32_t *strand_code1,
                          const int64_t *row_of2, int64_t maxo2, const char *names2, const int64_t *name_off2,
                          const int32_t *strand_code2,
                          int32_t threads, char *err, int32_t errlen);

/*
 * icl_host_write_blocks with the headers assembled natively (wgdi/run_colliearity.py:143-151,194-195):
 *   "# Alignment <k>:" + the text " score=<score> pvalue=<ptext[pidx]> N=<members> <pair_text[pair]> plus|minus" cut at its
 *   first ':' (the reference splits the finished header at ':' again, which truncates it for a chromosome name with one);
 *   plus iff loc2 of the first gene pair < loc2 of the second.  ptext: the distinct p values as Python prints them
 *   (the caller formats those few), pair_text: "<chr1>&<chr2>" per pair id.
 */
int icl_host_write_collinearity(const char *path, int64_t n_sel, const int64_t *score,
                                const char *ptext, const int64_t *ptext_off, const int32_t *pidx,
                                const char *pair_text, const int64_t *pair_text_off, const int32_t *pair,
                                const int64_t *mem_begin, const int64_t *mem_end,
                                const int32_t *loc1, const int32_t *loc2,
                                const int32_t *chr1_code, const int32_t *chr2_code,
                                const int64_t *row_of1, int64_t maxo1, const char *names1, const int64_t *name_off1,
                                const int32_t *strand_code1,
                                const int64_t *row_of2, int64_t maxo2, const char *names2, const int64_t *name_off2,
                                const int32_t *strand_code2,
                                int32_t threads, char *err, int32_t errlen);

/*
 * `-bi` input: a WGDI collinearity file (as -icl writes it) back into arrays — wgdi/base.py:120-141 (read_collinearity) plus
 * what wgdi/block_info.py:18-75 derives per block member: the gene rows of the two names (-1: not in the gene table), the
 * two positions as float(), and "_".join(str(int(position))) per block.  Blocks: header tokens 0 and 2 of
 * re.findall(r'[\.\d+]+', header) (alignment number, p value) as text; a header without members is dropped like the
 * reference drops it.  Returns 1 WITHOUT reading when the file is not the plain case (a byte that is not printable ASCII,
 * ' ' or '\n', a position that is not a plain integer, a short line): the caller then uses its Python reader, which behaves
 * — and fails — like the reference.
 */
typedef struct icl_collinearity {
    int64_t n_blocks, n_members;
    char *id_text; int64_t *id_off;             /* [n_blocks + 1] */
    char *pv_text; int64_t *pv_off;
    int64_t *off;                               /* [n_blocks + 1] members of block b: off[b] .. off[b + 1] */
    int32_t *g1, *g2;                           /* [n_members] */
    double *loc1, *loc2;                        /* [n_members] */
    char *block1_text; int64_t *block1_off;     /* [n_blocks + 1] */
    char *block2_text; int64_t *block2_off;
} icl_collinearity;
int icl_host_read_collinearity(const char *path, const char *ids1, const int64_t *off1, int64_t n1,
                               const char *ids2, const int64_t *off2, int64_t n2,
                               icl_collinearity *out, char *err, int32_t errlen);
void icl_host_free_collinearity(icl_collinearity *c);

/*
 * One comparison sharded over several GPUs (the pairs of wgdi/run_colliearity.py:123-136 over ranks instead of worker
 * processes): write one rank's blocks straight to their place in the merged, emission-ordered arrays — what the reference's
 * parent process does when it concatenates the workers' texts in chunk order (:133-141).  A rank's blocks are pair-major, so
 * a pair's run of blocks is contiguous on both sides:
 *   block b of local pair q, global pair g = global_pair[q]:  destination block_base[g] + (b - first block of the run),
 *   its gene pairs from member_base[g] + (member_off[b] - member_off[first block of the run]).
 * block_base / member_base: exclusive sums over the per-pair counts of ALL ranks.  Per-block arrays (n_block_fields of
 * block_bytes[k] bytes per element) and per-gene-pair arrays (n_member_fields) are copied as they are; out_pair receives
 * the global pair ids, out_member_off the new member offsets, and member_idx (may be NULL) is mapped through `rows`
 * (the rank's upload index -> the comparison's) into out_member_idx.  wgdi_b200/shard.py:place_blocks is the same in numpy.
 */
int icl_host_place_blocks(int64_t n_blocks, const int32_t *pair, const int32_t *global_pair,
                          const int64_t *block_base, const int64_t *member_base, const int64_t *member_off,
                          int32_t n_block_fields, const void *const *block_src, void *const *block_dst, const int32_t *block_bytes,
                          int32_t n_member_fields, const void *const *member_src, void *const *member_dst, const int32_t *member_bytes,
                          int32_t *out_pair, int64_t *out_member_off,
                          const int32_t *member_idx, const int64_t *rows, int32_t *out_member_idx);

#ifdef __cplusplus
}
#endif
#endif /* ICL_HOST_H */
