// icl_blockinfo.cu — device part of `wgdi -bi` behind include/icl_blockinfo.h (sm_100a).
//
// One warp per block of the collinearity file.  A lane owns a member (gene pair) of the block and walks the kept
// BLAST hits of the member's query gene (at most repeat_number, contiguous in the CSR):
//   * the hit whose subject is the member's own partner gives the homo1..homo5 classes (block_info.py:44-45,109-116);
//   * every hit whose subject is one of the block's id2 genes (binary search in the block's sorted subject list) and
//     whose subject lies 1..repeat_number genes away from the member's partner is a row of the reference's merge
//     (tandem_ratio, block_info.py:122-134).
// Integer work only; the ratios and means are formed on the host exactly as the reference's Python does.
#include "../../include/icl_blockinfo.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <string>

namespace {

struct BiArgs {
    const long long *hit_off, *blk_off;
    const int *hit_g2, *hit_rank, *hit_order2, *mem_g1, *mem_g2, *mem_order2, *mem_g2_sorted;
    unsigned char *pass;          // per member: has at least one tandem row
    int *homo_sum, *homo_cnt, *tandem_uniq;
    long long *tandem_rows;
    long long n_blocks;
    int rn;
};

__device__ __forceinline__ bool in_sorted(const int *a, long long lo, long long hi, int v) {
    while (lo < hi) {
        long long mid = lo + ((hi - lo) >> 1);
        int x = a[mid];
        if (x == v) return true;
        if (x < v) lo = mid + 1; else hi = mid;
    }
    return false;
}

__global__ void __launch_bounds__(128) k_block_stats(BiArgs A) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= A.n_blocks) return;
    const long long m0 = A.blk_off[b], m1 = A.blk_off[b + 1];
    int hs[5] = {0, 0, 0, 0, 0};
    int cnt = 0, uniq = 0;
    long long rows = 0;
    for (long long j = m0 + lane; j < m1; j += 32) {
        const int g1 = A.mem_g1[j], g2 = A.mem_g2[j];
        bool pass = false;
        if (g1 >= 0) {
            const int o2 = g2 >= 0 ? A.mem_order2[j] : 0;
            for (long long h = A.hit_off[g1], h1 = A.hit_off[g1 + 1]; h < h1; h++) {
                const int t = A.hit_g2[h];
                if (t == g2) {                                    // the member's own pair is a kept hit
                    const int r = A.hit_rank[h];
#pragma unroll
                    for (int i = 1; i <= 5; i++) hs[i - 1] += r < i ? 1 : (r < i + 5 ? 0 : -1);
                    cnt++;
                }
                if (g2 >= 0 && in_sorted(A.mem_g2_sorted, m0, m1, t)) {
                    int d = A.hit_order2[h] - o2;
                    d = d < 0 ? -d : d;
                    if (d > 0 && d <= A.rn) { rows++; pass = true; }
                }
            }
        }
        A.pass[j] = pass ? 1 : 0;
    }
    __syncwarp();
    // distinct id1 among the members that produced rows: count a member unless an earlier one has the same id1
    for (long long j = m0 + lane; j < m1; j += 32) {
        if (!A.pass[j]) continue;
        const int g1 = A.mem_g1[j];
        bool first = true;
        for (long long k = m0; k < j && first; k++)
            if (A.mem_g1[k] == g1 && A.pass[k]) first = false;
        uniq += first ? 1 : 0;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) hs[i] = __reduce_add_sync(0xffffffffu, hs[i]);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    uniq = __reduce_add_sync(0xffffffffu, uniq);
    for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
    if (lane == 0) {
        for (int i = 0; i < 5; i++) A.homo_sum[5 * b + i] = hs[i];
        A.homo_cnt[b] = cnt;
        A.tandem_rows[b] = rows;
        A.tandem_uniq[b] = uniq;
    }
}

struct DevBufs {
    void *p[16];
    int n = 0;
    ~DevBufs() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};

int fail(char *err, int32_t errlen, int code, const std::string &msg) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg.c_str());
    return code;
}

}  // namespace

extern "C" const char *icl_bi_backend(void) { return "cuda"; }

extern "C" int icl_bi_block_stats(int device, const icl_bi_input *in, int32_t repeat_number,
                                  int32_t *homo_sum, int32_t *homo_cnt, int64_t *tandem_rows, int32_t *tandem_uniq,
                                  char *err, int32_t errlen) {
    if (!in || !homo_sum || !homo_cnt || !tandem_rows || !tandem_uniq || !in->hit_off || !in->blk_off || in->n_blocks < 0 ||
        in->n_genes1 < 0 || repeat_number < 0)
        return fail(err, errlen, -1, "icl_bi_block_stats: bad argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(err, errlen, -2, std::string("icl_bi_block_stats: ") + cudaGetErrorString(e));
    const int64_t B = in->n_blocks, NH = in->hit_off[in->n_genes1], NM = in->blk_off[B];
    if (B == 0) return 0;
    if (NM > 0 && (!in->mem_g1 || !in->mem_g2 || !in->mem_order2 || !in->mem_g2_sorted))
        return fail(err, errlen, -1, "icl_bi_block_stats: null member array");
    if (NH > 0 && (!in->hit_g2 || !in->hit_rank || !in->hit_order2))
        return fail(err, errlen, -1, "icl_bi_block_stats: null hit array");
    DevBufs d;
    cudaStream_t s = nullptr;
#define BI_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { if (s) cudaStreamDestroy(s); \
        return fail(err, errlen, -2, std::string("icl_bi_block_stats: ") + cudaGetErrorString(e_)); } } while (0)
    BI_CUDA(cudaStreamCreate(&s));
    auto up = [&](const void *h, size_t bytes, void **out) -> cudaError_t {
        cudaError_t r = cudaMalloc(out, bytes ? bytes : 16);
        if (r != cudaSuccess) return r;
        d.p[d.n++] = *out;
        return bytes ? cudaMemcpyAsync(*out, h, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    BiArgs A{};
    void *q = nullptr;
    BI_CUDA(up(in->hit_off, (size_t)(in->n_genes1 + 1) * 8, &q)); A.hit_off = (const long long *)q;
    BI_CUDA(up(in->blk_off, (size_t)(B + 1) * 8, &q)); A.blk_off = (const long long *)q;
    BI_CUDA(up(in->hit_g2, (size_t)NH * 4, &q)); A.hit_g2 = (const int *)q;
    BI_CUDA(up(in->hit_rank, (size_t)NH * 4, &q)); A.hit_rank = (const int *)q;
    BI_CUDA(up(in->hit_order2, (size_t)NH * 4, &q)); A.hit_order2 = (const int *)q;
    BI_CUDA(up(in->mem_g1, (size_t)NM * 4, &q)); A.mem_g1 = (const int *)q;
    BI_CUDA(up(in->mem_g2, (size_t)NM * 4, &q)); A.mem_g2 = (const int *)q;
    BI_CUDA(up(in->mem_order2, (size_t)NM * 4, &q)); A.mem_order2 = (const int *)q;
    BI_CUDA(up(in->mem_g2_sorted, (size_t)NM * 4, &q)); A.mem_g2_sorted = (const int *)q;
    void *o_pass, *o_sum, *o_cnt, *o_rows, *o_uniq;
    BI_CUDA(cudaMalloc(&o_pass, (size_t)(NM ? NM : 1))); d.p[d.n++] = o_pass;
    BI_CUDA(cudaMalloc(&o_sum, (size_t)B * 5 * 4)); d.p[d.n++] = o_sum;
    BI_CUDA(cudaMalloc(&o_cnt, (size_t)B * 4)); d.p[d.n++] = o_cnt;
    BI_CUDA(cudaMalloc(&o_rows, (size_t)B * 8)); d.p[d.n++] = o_rows;
    BI_CUDA(cudaMalloc(&o_uniq, (size_t)B * 4)); d.p[d.n++] = o_uniq;
    A.pass = (unsigned char *)o_pass; A.homo_sum = (int *)o_sum; A.homo_cnt = (int *)o_cnt;
    A.tandem_rows = (long long *)o_rows; A.tandem_uniq = (int *)o_uniq;
    A.n_blocks = B; A.rn = repeat_number;
    const unsigned grid = (unsigned)((B + 3) / 4);                          // 4 warps (blocks of the file) per CTA
    k_block_stats<<<grid, 128, 0, s>>>(A);
    BI_CUDA(cudaGetLastError());
    BI_CUDA(cudaMemcpyAsync(homo_sum, o_sum, (size_t)B * 5 * 4, cudaMemcpyDeviceToHost, s));
    BI_CUDA(cudaMemcpyAsync(homo_cnt, o_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    BI_CUDA(cudaMemcpyAsync(tandem_rows, o_rows, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    BI_CUDA(cudaMemcpyAsync(tandem_uniq, o_uniq, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    BI_CUDA(cudaStreamSynchronize(s));
    cudaStreamDestroy(s);
    return 0;
}
