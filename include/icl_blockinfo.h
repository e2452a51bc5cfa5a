/*
 * icl_blockinfo.h — C ABI of the device part of `wgdi -bi` (block information).
 *
 * The reference (wgdi/block_info.py) walks every block of a collinearity file in Python and, per block, does one
 * pandas lookup per gene pair (homo1..homo5, block_info.py:44-45 over the table of blast_homo :102-120) and two
 * `isin` scans over the whole BLAST table plus a merge (tandem_ratio, :122-134).  Here all blocks are processed by
 * one kernel launch (one warp per block) over a CSR of the kept BLAST hits; the host side (wgdi_b200/block_info.py)
 * parses the files, ranks the hits and writes the CSV.
 *
 * All arrays are host memory, caller-owned, read during the call only.  Returns 0 or a negative code and fills
 * `err`.  There is no CPU fallback: without a usable CUDA device the call fails.
 */
#ifndef ICL_BLOCKINFO_H
#define ICL_BLOCKINFO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct icl_bi_input {
    /* BLAST hits kept by blast_homo (block_info.py:104-105: per query the first repeat_number hits by bit score),
       grouped by query gene: hits of query g are hit_off[g] .. hit_off[g+1] */
    int64_t n_genes1;
    const int64_t *hit_off;      /* [n_genes1 + 1] */
    const int32_t *hit_g2;       /* subject gene (row of the genome-2 gene table) */
    const int32_t *hit_rank;     /* 0-based rank of the hit among its query's hits (block_info.py:104) */
    const int32_t *hit_order2;   /* gene order of the subject (`chr2_order`, block_info.py:119) */
    /* blocks of the collinearity file: members of block b are blk_off[b] .. blk_off[b+1] */
    int64_t n_blocks;
    const int64_t *blk_off;      /* [n_blocks + 1] */
    const int32_t *mem_g1;       /* gene-table row of the member's id1, -1 when the name is unknown */
    const int32_t *mem_g2;       /* same for id2 */
    const int32_t *mem_order2;   /* gene order of id2 (`order2`, block_info.py:125); ignored where mem_g2 < 0 */
    const int32_t *mem_g2_sorted;/* per block: the mem_g2 values of the block in ascending order */
} icl_bi_input;

/*
 * Per block b:
 *   homo_sum[5*b + i-1] = sum over the members whose (id1, id2) is a kept hit of  homo_i(rank)
 *                         homo_i = 1 if rank < i, 0 if rank < i + 5, else -1          (block_info.py:109-116)
 *   homo_cnt[b]         = number of such members (the reference averages over them, :98-100)
 *   tandem_rows[b]      = rows of the merge of :128-133: pairs (member m, kept hit h of m's id1) with h's subject
 *                         among the block's id2 genes and 0 < |order2(h) - order2(m)| <= repeat_number
 *   tandem_uniq[b]      = distinct id1 among those rows (:134)
 */
int icl_bi_block_stats(int device, const icl_bi_input *in, int32_t repeat_number,
                       int32_t *homo_sum, int32_t *homo_cnt, int64_t *tandem_rows, int32_t *tandem_uniq,
                       char *err, int32_t errlen);

/* "cuda" */
const char *icl_bi_backend(void);

#ifdef __cplusplus
}
#endif
#endif /* ICL_BLOCKINFO_H */
