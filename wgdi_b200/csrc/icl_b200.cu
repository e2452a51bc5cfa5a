// icl_b200.cu — B200-native engine for the `wgdi -icl` hot path behind the C ABI of include/icl_b200.h.
//
// What the reference does per chromosome pair in Python/pandas (wgdi/collinearity.py:30-153, driven by
// wgdi/run_colliearity.py:158-170) is done here for ALL pairs of a comparison at once:
//
//   icl_upload   points (pair, loc1, loc2, grading) -> HBM
//   icl_dp       radix sort by (pair, loc1, loc2)            [CUB]        run_colliearity.py:164-166
//                dense row index + mirrored copy for the anti-diagonal direction
//                candidate ("edge") enumeration, fully parallel           collinearity.py:76-79 / :98-101
//                chaining DP, order-exact                                 collinearity.py:81-93 / :103-115
//   icl_extract  greedy block extraction in the caller's rank order       collinearity.py:49-68, 117-144
//                times / p-value pass in fp64                             collinearity.py:146-153
//
// Data layout in HBM (SoA, int32): the sorted points of all pairs are concatenated ("forward space",
// n points).  The anti-diagonal DP of the reference is the forward DP under loc1 -> -loc1, so a mirrored
// copy of every pair is appended ("V space" = 2n virtual points: [0,n) forward, [n,2n) mirrored) and one
// set of kernels serves both directions.  A work item is (pair, direction); item it = dir * P + pair.
//
// See DESIGN.md for the kernel inventory, the roofline of each kernel and the parity argument.
#include "../../include/icl_b200.h"
#include "icl_platform.h"

#include <algorithm>
#include <numeric>
#include <vector>

using namespace icl;

#ifndef ICL_EMU
#define ICL_LAMBDA [=] __host__ __device__
#else
#define ICL_LAMBDA [=]
#endif

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct icl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Status st;
    int64_t n = 0;      // points
    int32_t P = 0;      // chromosome pairs
    bool uploaded = false, dp_done = false;
    icl_params prm{};
    int bits1 = 1, bits2 = 1, bitsp = 1;

    // inputs in upload order
    Buf<int32_t> in_pair, in_loc1, in_loc2, in_grad, stats;
    // sort
    Buf<uint64_t> key_a, key_b;
    Buf<uint32_t> idx_a, perm;
    Buf<uint8_t> tmp;
    // layout
    Buf<int64_t> d_pair_off, d_rb;
    Buf<int32_t> d_rmin, d_nrows, d_item_order, d_pair_order;
    std::vector<int64_t> h_pair_off, h_rb;
    std::vector<int32_t> h_rmin, h_nrows, h_item_order, h_pair_order;
    int64_t R = 0;      // dense rows incl. one sentinel per pair (forward); the mirrored index follows at +R
    Buf<int32_t> loc1s, loc2s, pair_s, rowptr, X1, X2, G, itemof, mmap, imap;
    // edges
    Buf<uint32_t> ecnt, eoff;
    Buf<uint2> edges;
    int64_t n_edges = 0;
    // DP state (V space)
    Buf<int32_t> S, pred, depth;
    Buf<uint8_t> used;
    Buf<int64_t> score64;
    // extraction
    Buf<int32_t> order_in, tops, members, blk_top, blk_m, blk_L, blk_ms, blk_N, blk_L1, blk_L2, nblk, times;
    Buf<int64_t> blk_N1, d_out_base;
    Buf<double> blk_a;
    Buf<uint8_t> present;
    // gather
    Buf<int32_t> o_pair, o_dir, o_m, o_L, o_N, o_L1, o_L2, o_loc1, o_loc2, o_idx, o_slot;
    Buf<int64_t> o_score, o_N1;
    Buf<double> o_a;
    Buf<uint32_t> o_mcnt, o_moff;

    Timer timers[ICL_T_COUNT];
    int64_t counters[4] = {0, 0, 0, 0};
};

static std::string g_create_error;

static int fail(icl_ctx *c, int code, const std::string &msg) {
    c->st.code = code;
    c->st.msg = msg;
    return code;
}

// first index in [lo, hi) with a[idx] > v
template <class T>
ICL_HD int32_t upper_bound_i32(const T *a, int32_t lo, int32_t hi, T v) {
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] > v) hi = mid; else lo = mid + 1;
    }
    return lo;
}

static int bits_for(int64_t maxval) {
    int b = 1;
    while (b < 63 && (1ll << b) <= maxval) b++;
    return b;
}

// ------------------------------------------------------------------------------------------------
// input statistics (max loc1 / loc2, validity) — one warp-reduced atomic per warp
// ------------------------------------------------------------------------------------------------
#ifndef ICL_EMU
__global__ void __launch_bounds__(256) k_input_stats(int64_t n, int32_t P, const int32_t *pair, const int32_t *l1,
                                                     const int32_t *l2, int32_t *stats) {
    int32_t m1 = 0, m2 = 0, bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t a = l1[i], b = l2[i], p = pair[i];
        bad |= (a < 0) | (b < 0) | (p < 0) | (p >= P);
        m1 = max(m1, a);
        m2 = max(m2, b);
    }
    m1 = __reduce_max_sync(0xffffffffu, m1);
    m2 = __reduce_max_sync(0xffffffffu, m2);
    bad = __reduce_or_sync(0xffffffffu, (unsigned)bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&stats[0], m1);
        atomicMax(&stats[1], m2);
        if (bad) atomicOr(&stats[2], 1);
    }
}
#endif

static int input_stats(icl_ctx *c) {
    Status &st = c->st;
    ICL_CUDA(st, dev_memset(c->stats.p, 0, 4 * sizeof(int32_t), c->stream));
#ifndef ICL_EMU
    if (c->n > 0) {
        int64_t blocks = std::min<int64_t>((c->n + 255) / 256, 148 * 8);
        k_input_stats<<<(unsigned)blocks, 256, 0, c->stream>>>(c->n, c->P, c->in_pair.p, c->in_loc1.p, c->in_loc2.p, c->stats.p);
        ICL_CUDA(st, cudaGetLastError());
    }
#else
    for (int64_t i = 0; i < c->n; i++) {
        int32_t a = c->in_loc1.p[i], b = c->in_loc2.p[i], p = c->in_pair.p[i];
        if (a < 0 || b < 0 || p < 0 || p >= c->P) c->stats.p[2] = 1;
        c->stats.p[0] = std::max(c->stats.p[0], a);
        c->stats.p[1] = std::max(c->stats.p[1], b);
    }
#endif
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: lifetime
// ------------------------------------------------------------------------------------------------
extern "C" const char *icl_backend(void) {
#ifndef ICL_EMU
    return "cuda";
#else
    return "emu";
#endif
}

extern "C" int icl_create(int device, icl_ctx **out) {
    if (!out) { g_create_error = "icl_create: ctx pointer is NULL"; return ICL_ERR_ARG; }
    *out = nullptr;
#ifndef ICL_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        g_create_error = std::string("icl_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return ICL_ERR_CUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "icl_create: device index out of range"; return ICL_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return ICL_ERR_CUDA; }
#endif
    icl_ctx *c = new (std::nothrow) icl_ctx();
    if (!c) { g_create_error = "icl_create: out of host memory"; return ICL_ERR_NOMEM; }
    c->device = device;
#ifndef ICL_EMU
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete c; return ICL_ERR_CUDA; }
#endif
    if (c->stats.reserve(c->st, 4)) { g_create_error = c->st.msg; delete c; return ICL_ERR_CUDA; }
    *out = c;
    return ICL_OK;
}

extern "C" void icl_destroy(icl_ctx *c) {
    if (!c) return;
#ifndef ICL_EMU
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
#endif
    c->in_pair.release(); c->in_loc1.release(); c->in_loc2.release(); c->in_grad.release(); c->stats.release();
    c->key_a.release(); c->key_b.release(); c->idx_a.release(); c->perm.release(); c->tmp.release();
    c->d_pair_off.release(); c->d_rb.release(); c->d_rmin.release(); c->d_nrows.release();
    c->d_item_order.release(); c->d_pair_order.release();
    c->loc1s.release(); c->loc2s.release(); c->pair_s.release(); c->rowptr.release();
    c->X1.release(); c->X2.release(); c->G.release(); c->itemof.release(); c->mmap.release(); c->imap.release();
    c->ecnt.release(); c->eoff.release(); c->edges.release();
    c->S.release(); c->pred.release(); c->depth.release(); c->used.release(); c->score64.release();
    c->order_in.release(); c->tops.release(); c->members.release(); c->blk_top.release(); c->blk_m.release();
    c->blk_L.release(); c->blk_ms.release(); c->blk_N.release(); c->blk_L1.release(); c->blk_L2.release();
    c->nblk.release(); c->times.release(); c->blk_N1.release(); c->d_out_base.release(); c->blk_a.release();
    c->present.release();
    c->o_pair.release(); c->o_dir.release(); c->o_m.release(); c->o_L.release(); c->o_N.release(); c->o_L1.release();
    c->o_L2.release(); c->o_loc1.release(); c->o_loc2.release(); c->o_idx.release(); c->o_slot.release();
    c->o_score.release(); c->o_N1.release(); c->o_a.release(); c->o_mcnt.release(); c->o_moff.release();
    for (auto &t : c->timers) t.destroy();
#ifndef ICL_EMU
    cudaStreamDestroy(c->stream);
#endif
    delete c;
}

extern "C" const char *icl_last_error(const icl_ctx *c) {
    return c ? c->st.msg.c_str() : g_create_error.c_str();
}

extern "C" int icl_host_alloc(icl_ctx *c, int64_t bytes, void **ptr) {
    if (!c || !ptr || bytes < 0) return ICL_ERR_ARG;
#ifndef ICL_EMU
    ICL_CUDA(c->st, cudaSetDevice(c->device));
    ICL_CUDA(c->st, cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 16), cudaHostAllocDefault));
#else
    *ptr = malloc((size_t)(bytes ? bytes : 16));
    if (!*ptr) return fail(c, ICL_ERR_NOMEM, "icl_host_alloc: out of memory");
#endif
    return ICL_OK;
}

extern "C" int icl_host_free(icl_ctx *c, void *ptr) {
    if (!c) return ICL_ERR_ARG;
#ifndef ICL_EMU
    ICL_CUDA(c->st, cudaFreeHost(ptr));
#else
    free(ptr);
#endif
    return ICL_OK;
}

// ------------------------------------------------------------------------------------------------
// ABI: upload
// ------------------------------------------------------------------------------------------------
extern "C" int icl_upload(icl_ctx *c, int64_t n, int32_t P, const int32_t *pair_id, const int32_t *loc1,
                          const int32_t *loc2, const int32_t *grading) {
    if (!c) return ICL_ERR_ARG;
    c->uploaded = c->dp_done = false;
    if (n < 0 || P < 0 || (n > 0 && (!pair_id || !loc1 || !loc2 || !grading)) || (n > 0 && P == 0))
        return fail(c, ICL_ERR_ARG, "icl_upload: bad argument");
    if (n >= (1ll << 30)) return fail(c, ICL_ERR_LIMIT, "icl_upload: n_points must be < 2^30");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    c->n = n;
    c->P = P;
    if (c->in_pair.reserve(st, n) || c->in_loc1.reserve(st, n) || c->in_loc2.reserve(st, n) || c->in_grad.reserve(st, n))
        return st.code;
    size_t b = (size_t)n * sizeof(int32_t);
    ICL_CUDA(st, h2d(c->in_pair.p, pair_id, b, c->stream));
    ICL_CUDA(st, h2d(c->in_loc1.p, loc1, b, c->stream));
    ICL_CUDA(st, h2d(c->in_loc2.p, loc2, b, c->stream));
    ICL_CUDA(st, h2d(c->in_grad.p, grading, b, c->stream));
    if (input_stats(c)) return st.code;
    int32_t hs[4] = {0, 0, 0, 0};
    ICL_CUDA(st, d2h(hs, c->stats.p, sizeof(hs), c->stream));
    ICL_CUDA(st, dev_sync(c->stream));
    if (hs[2]) return fail(c, ICL_ERR_ARG, "icl_upload: negative coordinate or pair id outside [0, n_pairs)");
    if (hs[0] >= (1 << 26) || hs[1] >= (1 << 26)) return fail(c, ICL_ERR_LIMIT, "icl_upload: coordinates must be < 2^26");
    c->bits1 = bits_for(hs[0]);
    c->bits2 = bits_for(hs[1]);
    c->bitsp = bits_for(P > 0 ? P - 1 : 0);
    if (c->bits1 + c->bits2 + c->bitsp > 64) return fail(c, ICL_ERR_LIMIT, "icl_upload: key does not fit 64 bits");
    c->uploaded = true;
    st.code = 0;
    return ICL_OK;
}

// ------------------------------------------------------------------------------------------------
// phase: sort + layout
// ------------------------------------------------------------------------------------------------
static int phase_sort(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->key_a.reserve(st, n) || c->key_b.reserve(st, n) || c->idx_a.reserve(st, n) || c->perm.reserve(st, n) ||
        c->loc1s.reserve(st, n) || c->loc2s.reserve(st, n) || c->pair_s.reserve(st, n) || c->G.reserve(st, 2 * n) ||
        c->d_pair_off.reserve(st, P + 1) || c->d_rmin.reserve(st, P + 1) || c->d_nrows.reserve(st, P + 1))
        return st.code;
    const int sh2 = c->bits2, shp = c->bits1 + c->bits2;
    {
        const int32_t *ip = c->in_pair.p, *i1 = c->in_loc1.p, *i2 = c->in_loc2.p;
        uint64_t *key = c->key_a.p;
        uint32_t *idx = c->idx_a.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            key[i] = ((uint64_t)(uint32_t)ip[i] << shp) | ((uint64_t)(uint32_t)i1[i] << sh2) | (uint64_t)(uint32_t)i2[i];
            idx[i] = (uint32_t)i;
        }));
    }
    // stable LSD radix sort: equal (pair, loc1, loc2) keep upload order (run_colliearity.py:164-166)
    if (n > 0 && radix_sort_pairs(st, c->tmp, c->key_a.p, c->key_b.p, c->idx_a.p, c->perm.p, n, shp + c->bitsp, s)) return st.code;
    {
        const uint64_t *key = c->key_b.p;
        const uint32_t *perm = c->perm.p;
        const int32_t *ig = c->in_grad.p;
        int32_t *l1 = c->loc1s.p, *l2 = c->loc2s.p, *ps = c->pair_s.p, *G = c->G.p;
        const uint64_t m1 = (1ull << c->bits1) - 1, m2 = (1ull << c->bits2) - 1;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            uint64_t k = key[i];
            l2[i] = (int32_t)(k & m2);
            l1[i] = (int32_t)((k >> sh2) & m1);
            ps[i] = (int32_t)(k >> shp);
            G[i] = ig[perm[i]];
        }));
        int64_t *off = c->d_pair_off.p;
        ICL_CUDA(st, foreach(s, (int64_t)P + 1, ICL_LAMBDA(int64_t p) {
            // first sorted position whose pair id is >= p
            int64_t lo = 0, hi = n;
            while (lo < hi) {
                int64_t mid = lo + ((hi - lo) >> 1);
                if ((int64_t)(key[mid] >> shp) >= p) hi = mid; else lo = mid + 1;
            }
            off[p] = lo;
        }));
        int32_t *rmin = c->d_rmin.p, *nrows = c->d_nrows.p;
        ICL_CUDA(st, foreach(s, (int64_t)P, ICL_LAMBDA(int64_t p) {
            int64_t a = off[p], b = off[p + 1];
            if (b > a) { rmin[p] = l1[a]; nrows[p] = l1[b - 1] - l1[a] + 1; }
            else { rmin[p] = 0; nrows[p] = 0; }
        }));
    }
    c->h_pair_off.assign(P + 1, 0);
    c->h_rmin.assign(P + 1, 0);
    c->h_nrows.assign(P + 1, 0);
    ICL_CUDA(st, d2h(c->h_pair_off.data(), c->d_pair_off.p, (size_t)(P + 1) * sizeof(int64_t), s));
    if (P > 0) {
        ICL_CUDA(st, d2h(c->h_rmin.data(), c->d_rmin.p, (size_t)P * sizeof(int32_t), s));
        ICL_CUDA(st, d2h(c->h_nrows.data(), c->d_nrows.p, (size_t)P * sizeof(int32_t), s));
    }
    ICL_CUDA(st, dev_sync(s));
    // dense row index: nrows[p] + 1 slots per pair; longest-processing-time order of items / pairs
    c->h_rb.assign(P + 1, 0);
    for (int32_t p = 0; p < P; p++) c->h_rb[p + 1] = c->h_rb[p] + c->h_nrows[p] + 1;
    c->R = c->h_rb[P];
    if (2 * c->R >= (1ll << 31)) return fail(c, ICL_ERR_LIMIT, "icl_dp: dense row index too large (sparse gene orders?)");
    c->h_pair_order.resize(P);
    std::iota(c->h_pair_order.begin(), c->h_pair_order.end(), 0);
    const std::vector<int64_t> &po = c->h_pair_off;
    std::stable_sort(c->h_pair_order.begin(), c->h_pair_order.end(),
                     [&](int32_t a, int32_t b) { return po[a + 1] - po[a] > po[b + 1] - po[b]; });
    c->h_item_order.resize(2 * (size_t)P);
    for (int32_t j = 0; j < P; j++) {
        c->h_item_order[2 * j] = c->h_pair_order[j];
        c->h_item_order[2 * j + 1] = P + c->h_pair_order[j];
    }
    if (c->d_rb.reserve(st, P + 1) || c->d_item_order.reserve(st, 2 * P + 1) || c->d_pair_order.reserve(st, P + 1)) return st.code;
    ICL_CUDA(st, h2d(c->d_rb.p, c->h_rb.data(), (size_t)(P + 1) * sizeof(int64_t), s));
    if (P > 0) {
        ICL_CUDA(st, h2d(c->d_item_order.p, c->h_item_order.data(), (size_t)2 * P * sizeof(int32_t), s));
        ICL_CUDA(st, h2d(c->d_pair_order.p, c->h_pair_order.data(), (size_t)P * sizeof(int32_t), s));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: dense row index, mirrored copy (V space)
// ------------------------------------------------------------------------------------------------
static int phase_setup(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, R = c->R;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->rowptr.reserve(st, 2 * R + 2) || c->X1.reserve(st, 2 * n) || c->X2.reserve(st, 2 * n) ||
        c->itemof.reserve(st, 2 * n) || c->mmap.reserve(st, n) || c->imap.reserve(st, n))
        return st.code;
    const int32_t *l1 = c->loc1s.p, *l2 = c->loc2s.p, *ps = c->pair_s.p;
    const int64_t *off = c->d_pair_off.p, *rb = c->d_rb.p;
    const int32_t *rmin = c->d_rmin.p, *nrows = c->d_nrows.p;
    int32_t *rowptr = c->rowptr.p;
    // forward row starts: the first point of a row fills the slots of every empty row before it
    ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
        int32_t p = ps[i];
        int64_t a = off[p], b = off[p + 1];
        int32_t row = l1[i] - rmin[p];
        int32_t prev = (i == a) ? -1 : l1[i - 1] - rmin[p];
        for (int32_t r = prev + 1; r <= row; r++) rowptr[rb[p] + r] = (int32_t)i;
        if (i == b - 1) rowptr[rb[p] + nrows[p]] = (int32_t)b;
    }));
    // pairs without points still own one sentinel slot
    ICL_CUDA(st, foreach(s, (int64_t)P, ICL_LAMBDA(int64_t p) {
        if (off[p + 1] == off[p]) rowptr[rb[p]] = (int32_t)off[p];
    }));
    int32_t *X1 = c->X1.p, *X2 = c->X2.p, *G = c->G.p, *itemof = c->itemof.p, *mmap = c->mmap.p, *imap = c->imap.p;
    ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
        int32_t p = ps[i];
        int64_t a = off[p], b = off[p + 1];
        int32_t row = l1[i] - rmin[p];
        int32_t rs = rowptr[rb[p] + row], re = rowptr[rb[p] + row + 1];
        int64_t v = n + a + (b - re) + (i - rs);     // rows descending, columns ascending
        mmap[i] = (int32_t)v;
        imap[v - n] = (int32_t)i;
        X1[i] = row;
        X2[i] = l2[i];
        itemof[i] = p;
        X1[v] = nrows[p] - 1 - row;
        X2[v] = l2[i];
        G[v] = G[i];
        itemof[v] = P + p;
    }));
    // mirrored row starts: mirrored row x is forward row nrows-1-x
    ICL_CUDA(st, foreach(s, R, ICL_LAMBDA(int64_t g) {
        int32_t lo = 0, hi = P;                       // last pair with rb[p] <= g
        while (lo < hi) {
            int32_t mid = lo + ((hi - lo + 1) >> 1);
            if (rb[mid] <= g) lo = mid; else hi = mid - 1;
        }
        int32_t p = lo;
        if (p >= P) return;
        int32_t x = (int32_t)(g - rb[p]);
        int64_t a = off[p], b = off[p + 1];
        rowptr[R + g] = (int32_t)(n + a + (b - rowptr[rb[p] + (nrows[p] - x)]));
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: candidate enumeration.  For source v the candidates are the points k with
//   0 < X1[k]-X1[v] < mg1  and  0 < X2[k]-X2[v] < mg2          (collinearity.py:76-79 / :98-101)
// visited row by row (nearest first), column ascending — the order pandas' boolean filter keeps.
// An edge stores what the sweep of collinearity.py:81-93 needs: target offset, dr, dc and
// s = grading_k - (dr + dc).  Candidates with s <= 0 can never update (collinearity.py:87) and only
// matter through the `break` test of :83-84, for which the last candidate of a row is sufficient
// (a break at an earlier s<=0 candidate of the row is re-taken at any later candidate of that row:
// dc grows, gap and row_old are unchanged).  Such candidates are therefore dropped unless row-last.
// ------------------------------------------------------------------------------------------------
struct EdgeEnum {
    const int32_t *X1, *X2, *G, *itemof, *rowptr, *nrows;
    const int64_t *rb;
    int64_t R;
    int32_t P, mg1, mg2;
    template <class Emit>
    ICL_HD void run(int64_t v, Emit emit) const {
        int32_t it = itemof[v];
        int32_t p = it >= P ? it - P : it;
        int64_t base = rb[p] + (it >= P ? R : 0);
        int32_t nr = nrows[p];
        int32_t x1 = X1[v], x2 = X2[v];
        for (int32_t dr = 1; dr < mg1 && x1 + dr < nr; dr++) {
            int32_t lo = rowptr[base + x1 + dr], hi = rowptr[base + x1 + dr + 1];
            if (lo == hi) continue;
            int32_t k = upper_bound_i32(X2, lo, hi, x2);
            for (; k < hi; k++) {
                int32_t dc = X2[k] - x2;
                if (dc >= mg2) break;
                int32_t s = G[k] - (dr + dc);
                bool last = (k + 1 == hi) || (X2[k + 1] - x2 >= mg2);
                if (s > 0 || last) emit(k, dr, dc, s > 0 ? s : 0);
            }
        }
    }
};

static int phase_edges(icl_ctx *c) {
    Status &st = c->st;
    const int64_t V = 2 * c->n;
    cudaStream_t s = c->stream;
    if (c->ecnt.reserve(st, V + 2) || c->eoff.reserve(st, V + 2)) return st.code;
    EdgeEnum ee{c->X1.p, c->X2.p, c->G.p, c->itemof.p, c->rowptr.p, c->d_nrows.p, c->d_rb.p, c->R, c->P, c->prm.mg1, c->prm.mg2};
    uint32_t *ecnt = c->ecnt.p;
    ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) {
        uint32_t cnt = 0;
        ee.run(v, [&](int32_t, int32_t, int32_t, int32_t) { cnt++; });
        ecnt[v] = cnt;
    }));
    if (exclusive_sum_u32(st, c->tmp, c->ecnt.p, c->eoff.p, V, s)) return st.code;
    uint32_t total = 0;
    ICL_CUDA(st, d2h(&total, c->eoff.p + V, sizeof(uint32_t), s));
    ICL_CUDA(st, dev_sync(s));
    c->n_edges = total;
    if (c->edges.reserve(st, (size_t)total + 1)) return st.code;
    const uint32_t *eoff = c->eoff.p;
    uint2 *edges = c->edges.p;
    ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) {
        uint32_t e = eoff[v];
        ee.run(v, [&](int32_t k, int32_t dr, int32_t dc, int32_t sc) {
            edges[e++] = make_uint2((uint32_t)(k - (int32_t)v), (uint32_t)dr | ((uint32_t)dc << 8) | ((uint32_t)sc << 16));
        });
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: chaining DP (collinearity.py:81-93 / :103-115), one work item = (pair, direction).
// v0: one thread walks the item's sources in order; the state machine is the reference's, verbatim:
//   gap = mg2, row_old = row; for each candidate: if dc > gap and row_k beyond row_old: break;
//   if s > 0 and score_k < score_i + s: update, gap = min(dc, gap), row_old = row_k.
// ------------------------------------------------------------------------------------------------
static int phase_chain(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->S.reserve(st, V) || c->pred.reserve(st, V) || c->depth.reserve(st, V) || c->used.reserve(st, V)) return st.code;
    ICL_CUDA(st, d2d(c->S.p, c->G.p, (size_t)V * sizeof(int32_t), s));            // get_matrix: score = grading
    ICL_CUDA(st, dev_memset(c->pred.p, 0xFF, (size_t)V * sizeof(int32_t), s));
    ICL_CUDA(st, dev_memset(c->used.p, 0, (size_t)V, s));
    int32_t *S = c->S.p, *pred = c->pred.p, *depth = c->depth.p;
    uint8_t *used = c->used.p;
    const int32_t *G = c->G.p, *order = c->d_item_order.p;
    const int64_t *off = c->d_pair_off.p;
    const uint32_t *eoff = c->eoff.p;
    const uint2 *edges = c->edges.p;
    const int32_t mg2 = c->prm.mg2;
    ICL_CUDA(st, foreach_warp(s, 2 * (int64_t)P, ICL_LAMBDA(int64_t j) {
        int32_t it = order[j];
        int32_t p = it >= P ? it - P : it;
        int64_t a = off[p] + (it >= P ? n : 0), b = off[p + 1] + (it >= P ? n : 0);
        for (int64_t v = a; v < b; v++) {
            int32_t Sv = S[v];
            int32_t dv = (Sv != G[v]) ? depth[v] : 1;
            int32_t gap = mg2, dr_old = 0;
            bool uv = false;
            for (uint32_t e = eoff[v], e1 = eoff[v + 1]; e < e1; e++) {
                uint2 rec = edges[e];
                int32_t dr = rec.y & 255, dc = (rec.y >> 8) & 255, sc = (int32_t)(rec.y >> 16);
                if (dc > gap && dr > dr_old) break;
                if (sc > 0) {
                    int64_t k = v + (int32_t)rec.x;
                    if (S[k] < Sv + sc) {
                        S[k] = Sv + sc;
                        pred[k] = (int32_t)v;
                        depth[k] = dv + 1;
                        used[k] = 1;
                        uv = true;
                        gap = dc < gap ? dc : gap;
                        dr_old = dr;
                    }
                }
            }
            if (uv) used[v] = 1;
        }
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: icl_dp
// ------------------------------------------------------------------------------------------------
extern "C" int icl_dp(icl_ctx *c, const icl_params *prm, int64_t *score_fwd, int64_t *score_rev) {
    if (!c) return ICL_ERR_ARG;
    if (!c->uploaded) return fail(c, ICL_ERR_STATE, "icl_dp: icl_upload has not succeeded");
    if (!prm || (c->n > 0 && (!score_fwd || !score_rev))) return fail(c, ICL_ERR_ARG, "icl_dp: null argument");
    if (prm->mg1 < 2 || prm->mg1 > 255 || prm->mg2 < 2 || prm->mg2 > 255)
        return fail(c, ICL_ERR_LIMIT, "icl_dp: mg values must be in 2..255");
    if (prm->g0 <= 0 || prm->over_gap < 1) return fail(c, ICL_ERR_ARG, "icl_dp: grading[0] and over_gap must be positive");
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    c->dp_done = false;
    c->prm = *prm;
    cudaStream_t s = c->stream;
    const int64_t n = c->n;
    c->timers[ICL_T_SORT].start(s);
    if (phase_sort(c)) return st.code;
    c->timers[ICL_T_SORT].stop(s);
    c->timers[ICL_T_SETUP].start(s);
    if (phase_setup(c)) return st.code;
    c->timers[ICL_T_SETUP].stop(s);
    c->timers[ICL_T_EDGES].start(s);
    if (phase_edges(c)) return st.code;
    c->timers[ICL_T_EDGES].stop(s);
    c->timers[ICL_T_CHAIN].start(s);
    if (phase_chain(c)) return st.code;
    c->timers[ICL_T_CHAIN].stop(s);
    c->timers[ICL_T_EXPORT].start(s);
    if (c->score64.reserve(st, 2 * n)) return st.code;
    {
        const int32_t *S = c->S.p, *mmap = c->mmap.p;
        int64_t *o = c->score64.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            o[i] = S[i];
            o[n + i] = S[mmap[i]];
        }));
    }
    if (n > 0) {
        ICL_CUDA(st, d2h(score_fwd, c->score64.p, (size_t)n * sizeof(int64_t), s));
        ICL_CUDA(st, d2h(score_rev, c->score64.p + n, (size_t)n * sizeof(int64_t), s));
    }
    c->timers[ICL_T_EXPORT].stop(s);
    ICL_CUDA(st, dev_sync(s));
    c->counters[0] = c->n_edges;
    c->counters[1] = 13;   // kernels launched by icl_dp (stats excluded): key, decode, pair_off, pair_meta, rowptr x2,
                           // mirror, rowptr_rev, edge count, edge fill, chain, export + CUB passes
    c->counters[3] = c->R;
    c->dp_done = true;
    return ICL_OK;
}

extern "C" int icl_layout(icl_ctx *c, int64_t *pair_off, int32_t *perm) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_done) return fail(c, ICL_ERR_STATE, "icl_layout: icl_dp has not succeeded");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    if (pair_off) memcpy(pair_off, c->h_pair_off.data(), (size_t)(c->P + 1) * sizeof(int64_t));
    if (perm && c->n > 0) {
        ICL_CUDA(st, d2h(perm, c->perm.p, (size_t)c->n * sizeof(int32_t), c->stream));
        ICL_CUDA(st, dev_sync(c->stream));
    }
    return ICL_OK;
}

extern "C" int icl_debug_dp(icl_ctx *c, int32_t dir, int32_t *score, int32_t *pred, int32_t *depth, uint8_t *used) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_done) return fail(c, ICL_ERR_STATE, "icl_debug_dp: icl_dp has not succeeded");
    if (dir != 0 && dir != 1) return fail(c, ICL_ERR_ARG, "icl_debug_dp: dir must be 0 or 1");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    const int64_t n = c->n;
    if (n == 0) return ICL_OK;
    cudaStream_t s = c->stream;
    Buf<int32_t> t1, t2, t3;
    Buf<uint8_t> t4;
    if (t1.reserve(st, n) || t2.reserve(st, n) || t3.reserve(st, n) || t4.reserve(st, n)) return st.code;
    {
        const int32_t *S = c->S.p, *G = c->G.p, *pr = c->pred.p, *dp = c->depth.p, *mmap = c->mmap.p, *imap = c->imap.p;
        const uint8_t *us = c->used.p;
        int32_t *o1 = t1.p, *o2 = t2.p, *o3 = t3.p;
        uint8_t *o4 = t4.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            int64_t v = dir ? mmap[i] : i;
            bool root = S[v] == G[v];
            o1[i] = S[v];
            int32_t q = root ? -1 : pr[v];
            o2[i] = (q < 0) ? -1 : (dir ? imap[q - n] : q);
            o3[i] = root ? 1 : dp[v];
            o4[i] = us[v];
        }));
    }
    if (score) ICL_CUDA(st, d2h(score, t1.p, (size_t)n * 4, s));
    if (pred) ICL_CUDA(st, d2h(pred, t2.p, (size_t)n * 4, s));
    if (depth) ICL_CUDA(st, d2h(depth, t3.p, (size_t)n * 4, s));
    if (used) ICL_CUDA(st, d2h(used, t4.p, (size_t)n, s));
    ICL_CUDA(st, dev_sync(s));
    t1.release(); t2.release(); t3.release(); t4.release();
    return ICL_OK;
}

// ------------------------------------------------------------------------------------------------
// phase: extraction (collinearity.py:49-68 with max_path :117-144).  Tops are visited in the caller's
// rank order; a top that is still present is either accepted with the present part of its path
// (L >= over_gap and m / L > coverage_ratio, :129) or removed alone (:143).
// The removed members of a path always form a root-side prefix (an ancestor has a strictly smaller
// score than its descendants, so it can only have been removed as part of an accepted block, which
// removes every present ancestor as well); hence the present members are the walk from the top up to
// the first removed node.
// ------------------------------------------------------------------------------------------------
static int phase_extract(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    const int64_t nslots = V / 3 + 2 * (int64_t)P + 2;
    if (c->tops.reserve(st, V) || c->members.reserve(st, V) || c->present.reserve(st, V) || c->nblk.reserve(st, 2 * P + 1) ||
        c->blk_top.reserve(st, nslots) || c->blk_m.reserve(st, nslots) || c->blk_L.reserve(st, nslots) ||
        c->blk_ms.reserve(st, nslots) || c->stats.reserve(st, 4))
        return st.code;
    ICL_CUDA(st, dev_memset(c->stats.p, 0, 4 * sizeof(int32_t), s));
    const int32_t *ps = c->pair_s.p, *mmap = c->mmap.p, *imap = c->imap.p, *ord = c->order_in.p;
    const int64_t *off = c->d_pair_off.p;
    int32_t *tops = c->tops.p, *stats = c->stats.p;
    ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
        int32_t p = ps[i];
        int64_t a = off[p], np_ = off[p + 1] - a;
        int32_t of = ord[i], orv = ord[n + i];
        if (of < 0 || of >= np_ || orv < 0 || orv >= np_) { atomic_or_i32(&stats[2], 1); of = 0; orv = 0; }
        tops[i] = (int32_t)(a + of);
        tops[n + i] = mmap[a + orv];
    }));
    ICL_CUDA(st, d2d(c->present.p, c->used.p, (size_t)V, s));      // drop usedtimes < 1 (:50 / :61)
    const int32_t *S = c->S.p, *G = c->G.p, *pred = c->pred.p, *depth = c->depth.p, *order = c->d_item_order.p;
    uint8_t *present = c->present.p;
    int32_t *members = c->members.p, *blk_top = c->blk_top.p, *blk_m = c->blk_m.p, *blk_L = c->blk_L.p, *blk_ms = c->blk_ms.p;
    int32_t *nblk = c->nblk.p;
    const int32_t over_gap = c->prm.over_gap;
    const double cov = c->prm.coverage_ratio;
    ICL_CUDA(st, foreach_warp(s, 2 * (int64_t)P, ICL_LAMBDA(int64_t j) {
        int32_t it = order[j];
        int32_t p = it >= P ? it - P : it;
        int64_t a = off[p] + (it >= P ? n : 0), b = off[p + 1] + (it >= P ? n : 0);
        int64_t slot = a / 3 + it;
        int64_t cursor = a;
        int32_t nb = 0;
        for (int64_t t = a; t < b; t++) {
            int32_t top = tops[t];
            if (!present[top]) continue;
            int32_t L = (S[top] != G[top]) ? depth[top] : 1;
            int32_t m = 0;
            int32_t q = top;
            while (true) {                                  // present part of the path
                m++;
                if (S[q] == G[q]) break;                    // root
                q = pred[q];
                if (!present[q]) break;
            }
            if (L >= over_gap && (double)m / (double)L > cov) {
                q = top;
                for (int32_t k = 0; k < m; k++) {           // members by loc1 ascending (:140)
                    present[q] = 0;
                    if (it >= P) members[cursor + k] = imap[q - n];
                    else members[cursor + (m - 1 - k)] = q;
                    q = pred[q];
                }
                blk_top[slot + nb] = top;
                blk_m[slot + nb] = m;
                blk_L[slot + nb] = L;
                blk_ms[slot + nb] = (int32_t)cursor;
                cursor += m;
                nb++;
            } else {
                present[top] = 0;
            }
        }
        nblk[it] = nb;
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: times / p value (collinearity.py:131-139, 146-153), one chromosome pair per worker because
// `times` is shared by the pair's two directions and by all of its blocks, in acceptance order.
// ------------------------------------------------------------------------------------------------
ICL_HD double icl_p_raw(int64_t score, int32_t m, int32_t g0, int64_t N1, int32_t N, int32_t L1, int32_t L2) {
#if defined(__CUDA_ARCH__)
    // collinearity.py:152, left to right, round-to-nearest binary64, no contraction
    double a = __dsub_rn(1.0, __ddiv_rn(__ddiv_rn((double)score, (double)m), (double)g0));
    a = __dmul_rn(a, (double)(N1 - m + 1));
    a = __ddiv_rn(a, (double)N);
    a = __dmul_rn(a, (double)(L1 - m + 1));
    a = __dmul_rn(a, (double)(L2 - m + 1));
    a = __ddiv_rn(a, (double)L1);
    a = __ddiv_rn(a, (double)L2);
    return a;
#else
    volatile double a = 1.0 - (double)score / (double)m / (double)g0;
    a = a * (double)(N1 - m + 1);
    a = a / (double)N;
    a = a * (double)(L1 - m + 1);
    a = a * (double)(L2 - m + 1);
    a = a / (double)L1;
    a = a / (double)L2;
    return a;
#endif
}

static int phase_pvalue(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    const int64_t nslots = V / 3 + 2 * (int64_t)P + 2;
    if (c->times.reserve(st, n) || c->blk_N.reserve(st, nslots) || c->blk_N1.reserve(st, nslots) ||
        c->blk_L1.reserve(st, nslots) || c->blk_L2.reserve(st, nslots) || c->blk_a.reserve(st, nslots))
        return st.code;
    int32_t *times = c->times.p;
    ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) { times[i] = 1; }));       // get_matrix :34
    const int32_t *l1 = c->loc1s.p, *l2 = c->loc2s.p, *S = c->S.p, *rowptr = c->rowptr.p, *rmin = c->d_rmin.p;
    const int32_t *members = c->members.p, *blk_top = c->blk_top.p, *blk_m = c->blk_m.p, *blk_ms = c->blk_ms.p;
    const int32_t *nblk = c->nblk.p, *porder = c->d_pair_order.p;
    const int64_t *off = c->d_pair_off.p, *rb = c->d_rb.p;
    int32_t *bN = c->blk_N.p, *bL1 = c->blk_L1.p, *bL2 = c->blk_L2.p;
    int64_t *bN1 = c->blk_N1.p;
    double *ba = c->blk_a.p;
    const int32_t g0 = c->prm.g0;
    ICL_CUDA(st, foreach_warp(s, (int64_t)P, ICL_LAMBDA(int64_t j) {
        int32_t p = porder[j];
        for (int32_t d = 0; d < 2; d++) {
            int32_t it = d * P + p;
            int64_t a = off[p] + (d ? n : 0);
            int64_t slot = a / 3 + it;
            for (int32_t bi = 0; bi < nblk[it]; bi++) {
                int32_t m = blk_m[slot + bi], ms = blk_ms[slot + bi];
                int32_t l1min = l1[members[ms]], l1max = l1[members[ms + m - 1]];     // members are loc1-ascending
                int32_t l2min = l2[members[ms]], l2max = l2min;
                for (int32_t k = 1; k < m; k++) {
                    int32_t y = l2[members[ms + k]];
                    l2min = y < l2min ? y : l2min;
                    l2max = y > l2max ? y : l2max;
                }
                int32_t j0 = rowptr[rb[p] + (l1min - rmin[p])], j1 = rowptr[rb[p] + (l1max - rmin[p]) + 1];
                int32_t N = 0;
                int64_t N1 = 0;
                for (int32_t q = j0; q < j1; q++) {                                      // :134-137 inclusive box
                    int32_t y = l2[q];
                    if (y >= l2min && y <= l2max) { N1 += times[q]; N++; times[q]++; }   // :148-150
                }
                int32_t L1 = l1max - l1min + 1, L2 = l2max - l2min + 1;
                bN[slot + bi] = N;
                bN1[slot + bi] = N1;
                bL1[slot + bi] = L1;
                bL2[slot + bi] = L2;
                ba[slot + bi] = icl_p_raw((int64_t)S[blk_top[slot + bi]], m, g0, N1, N, L1, L2);
            }
        }
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: icl_extract
// ------------------------------------------------------------------------------------------------
template <class T>
static T *host_array(int64_t n) { return (T *)malloc((size_t)(n > 0 ? n : 1) * sizeof(T)); }

extern "C" void icl_free_blocks(icl_blocks *b) {
    if (!b) return;
    free(b->pair); free(b->dir); free(b->score); free(b->m); free(b->L); free(b->N); free(b->N1);
    free(b->L1); free(b->L2); free(b->p_raw); free(b->member_off); free(b->loc1); free(b->loc2); free(b->member_idx);
    memset(b, 0, sizeof(*b));
}

extern "C" int icl_extract(icl_ctx *c, const int32_t *order_fwd, const int32_t *order_rev, icl_blocks *out) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_done) return fail(c, ICL_ERR_STATE, "icl_extract: icl_dp has not succeeded");
    if (!out || (c->n > 0 && (!order_fwd || !order_rev))) return fail(c, ICL_ERR_ARG, "icl_extract: null argument");
    memset(out, 0, sizeof(*out));
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    cudaStream_t s = c->stream;
    const int64_t n = c->n;
    const int32_t P = c->P;
    c->timers[ICL_T_EXTRACT].start(s);
    if (c->order_in.reserve(st, 2 * n)) return st.code;
    if (n > 0) {
        ICL_CUDA(st, h2d(c->order_in.p, order_fwd, (size_t)n * sizeof(int32_t), s));
        ICL_CUDA(st, h2d(c->order_in.p + n, order_rev, (size_t)n * sizeof(int32_t), s));
    }
    if (phase_extract(c)) return st.code;
    c->timers[ICL_T_EXTRACT].stop(s);
    c->timers[ICL_T_PVALUE].start(s);
    if (phase_pvalue(c)) return st.code;
    c->timers[ICL_T_PVALUE].stop(s);
    c->timers[ICL_T_GATHER].start(s);
    // per-item block counts -> output offsets in (pair, direction) order
    std::vector<int32_t> h_nblk(2 * (size_t)P + 1, 0);
    int32_t hs[4] = {0, 0, 0, 0};
    if (P > 0) ICL_CUDA(st, d2h(h_nblk.data(), c->nblk.p, (size_t)2 * P * sizeof(int32_t), s));
    ICL_CUDA(st, d2h(hs, c->stats.p, sizeof(hs), s));
    ICL_CUDA(st, dev_sync(s));
    if (hs[2]) return fail(c, ICL_ERR_ARG, "icl_extract: order array holds an index outside its pair segment");
    std::vector<int64_t> h_out_base(2 * (size_t)P + 1, 0);      // output rank r = 2*pair + dir
    for (int32_t r = 0; r < 2 * P; r++) {
        int32_t it = (r & 1) * P + (r >> 1);
        h_out_base[r + 1] = h_out_base[r] + h_nblk[it];
    }
    const int64_t NB = h_out_base[2 * (size_t)P];
    if (c->d_out_base.reserve(st, 2 * P + 1)) return st.code;
    ICL_CUDA(st, h2d(c->d_out_base.p, h_out_base.data(), (size_t)(2 * P + 1) * sizeof(int64_t), s));
    if (c->o_pair.reserve(st, NB) || c->o_dir.reserve(st, NB) || c->o_m.reserve(st, NB) || c->o_L.reserve(st, NB) ||
        c->o_N.reserve(st, NB) || c->o_L1.reserve(st, NB) || c->o_L2.reserve(st, NB) || c->o_slot.reserve(st, NB) ||
        c->o_score.reserve(st, NB) || c->o_N1.reserve(st, NB) || c->o_a.reserve(st, NB) ||
        c->o_mcnt.reserve(st, NB + 2) || c->o_moff.reserve(st, NB + 2))
        return st.code;
    {
        const int64_t *ob = c->d_out_base.p, *off = c->d_pair_off.p;
        const int32_t *S = c->S.p, *blk_top = c->blk_top.p, *blk_m = c->blk_m.p, *blk_L = c->blk_L.p, *bN = c->blk_N.p;
        const int32_t *bL1 = c->blk_L1.p, *bL2 = c->blk_L2.p;
        const int64_t *bN1 = c->blk_N1.p;
        const double *ba = c->blk_a.p;
        int32_t *o_pair = c->o_pair.p, *o_dir = c->o_dir.p, *o_m = c->o_m.p, *o_L = c->o_L.p, *o_N = c->o_N.p;
        int32_t *o_L1 = c->o_L1.p, *o_L2 = c->o_L2.p, *o_slot = c->o_slot.p;
        int64_t *o_score = c->o_score.p, *o_N1 = c->o_N1.p;
        double *o_a = c->o_a.p;
        uint32_t *o_mcnt = c->o_mcnt.p;
        ICL_CUDA(st, foreach(s, NB, ICL_LAMBDA(int64_t g) {
            int32_t lo = 0, hi = 2 * P;                 // last rank with ob[r] <= g
            while (lo < hi) {
                int32_t mid = lo + ((hi - lo + 1) >> 1);
                if (ob[mid] <= g) lo = mid; else hi = mid - 1;
            }
            int32_t r = lo, p = r >> 1, d = r & 1, it = d * P + p;
            int64_t a = off[p] + (d ? n : 0);
            int64_t sl = a / 3 + it + (g - ob[r]);
            o_pair[g] = p;
            o_dir[g] = d;
            o_m[g] = blk_m[sl];
            o_mcnt[g] = (uint32_t)blk_m[sl];
            o_L[g] = blk_L[sl];
            o_N[g] = bN[sl];
            o_N1[g] = bN1[sl];
            o_L1[g] = bL1[sl];
            o_L2[g] = bL2[sl];
            o_a[g] = ba[sl];
            o_score[g] = S[blk_top[sl]];
            o_slot[g] = (int32_t)sl;
        }));
    }
    if (exclusive_sum_u32(st, c->tmp, c->o_mcnt.p, c->o_moff.p, NB, s)) return st.code;
    uint32_t NM = 0;
    ICL_CUDA(st, d2h(&NM, c->o_moff.p + NB, sizeof(uint32_t), s));
    ICL_CUDA(st, dev_sync(s));
    if (c->o_loc1.reserve(st, NM) || c->o_loc2.reserve(st, NM) || c->o_idx.reserve(st, NM)) return st.code;
    {
        const uint32_t *moff = c->o_moff.p, *perm = c->perm.p;
        const int32_t *o_slot = c->o_slot.p, *blk_ms = c->blk_ms.p, *members = c->members.p, *l1 = c->loc1s.p, *l2 = c->loc2s.p;
        int32_t *o_loc1 = c->o_loc1.p, *o_loc2 = c->o_loc2.p, *o_idx = c->o_idx.p;
        ICL_CUDA(st, foreach(s, (int64_t)NM, ICL_LAMBDA(int64_t g) {
            int64_t lo = 0, hi = NB;                    // last block with moff[b] <= g (blocks have m >= 1)
            while (lo < hi) {
                int64_t mid = lo + ((hi - lo + 1) >> 1);
                if ((int64_t)moff[mid] <= g) lo = mid; else hi = mid - 1;
            }
            int32_t i = members[blk_ms[o_slot[lo]] + (g - moff[lo])];
            o_loc1[g] = l1[i];
            o_loc2[g] = l2[i];
            o_idx[g] = (int32_t)perm[i];
        }));
    }
    out->n_blocks = NB;
    out->n_members = NM;
    out->pair = host_array<int32_t>(NB); out->dir = host_array<int32_t>(NB); out->score = host_array<int64_t>(NB);
    out->m = host_array<int32_t>(NB); out->L = host_array<int32_t>(NB); out->N = host_array<int32_t>(NB);
    out->N1 = host_array<int64_t>(NB); out->L1 = host_array<int32_t>(NB); out->L2 = host_array<int32_t>(NB);
    out->p_raw = host_array<double>(NB); out->member_off = host_array<int64_t>(NB + 1);
    out->loc1 = host_array<int32_t>(NM); out->loc2 = host_array<int32_t>(NM); out->member_idx = host_array<int32_t>(NM);
    if (!out->pair || !out->dir || !out->score || !out->m || !out->L || !out->N || !out->N1 || !out->L1 || !out->L2 ||
        !out->p_raw || !out->member_off || !out->loc1 || !out->loc2 || !out->member_idx) {
        icl_free_blocks(out);
        return fail(c, ICL_ERR_NOMEM, "icl_extract: out of host memory");
    }
    std::vector<uint32_t> h_moff((size_t)NB + 1, 0);
    if (NB > 0) {
        ICL_CUDA(st, d2h(out->pair, c->o_pair.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->dir, c->o_dir.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->score, c->o_score.p, (size_t)NB * 8, s));
        ICL_CUDA(st, d2h(out->m, c->o_m.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->L, c->o_L.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->N, c->o_N.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->N1, c->o_N1.p, (size_t)NB * 8, s));
        ICL_CUDA(st, d2h(out->L1, c->o_L1.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->L2, c->o_L2.p, (size_t)NB * 4, s));
        ICL_CUDA(st, d2h(out->p_raw, c->o_a.p, (size_t)NB * 8, s));
    }
    ICL_CUDA(st, d2h(h_moff.data(), c->o_moff.p, (size_t)(NB + 1) * 4, s));
    if (NM > 0) {
        ICL_CUDA(st, d2h(out->loc1, c->o_loc1.p, (size_t)NM * 4, s));
        ICL_CUDA(st, d2h(out->loc2, c->o_loc2.p, (size_t)NM * 4, s));
        ICL_CUDA(st, d2h(out->member_idx, c->o_idx.p, (size_t)NM * 4, s));
    }
    c->timers[ICL_T_GATHER].stop(s);
    ICL_CUDA(st, dev_sync(s));
    for (int64_t b = 0; b <= NB; b++) out->member_off[b] = h_moff[b];
    c->counters[2] = 7;   // tops, extract, times init, pvalue, gather blocks, gather members + CUB scan
    return ICL_OK;
}

extern "C" int icl_timings(icl_ctx *c, double *ms, int32_t n_slots) {
    if (!c || !ms) return ICL_ERR_ARG;
    int32_t k = n_slots < ICL_T_COUNT ? n_slots : ICL_T_COUNT;
    for (int32_t i = 0; i < k; i++) ms[i] = c->timers[i].ms();
    return k;
}

extern "C" int icl_counters(icl_ctx *c, int64_t *out, int32_t n_slots) {
    if (!c || !out) return ICL_ERR_ARG;
    int32_t k = n_slots < 4 ? n_slots : 4;
    for (int32_t i = 0; i < k; i++) out[i] = c->counters[i];
    return k;
}
