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
#include <cstdio>
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
// grading classes -> grading values of the current parameter set (icl_set_grading_table)
struct GradeTable {
    int32_t n = 0;                   // 0 = the uploaded values are the gradings themselves
    int32_t v[ICL_MAX_GRADE_CLASSES] = {};
};

struct icl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side streams: kernels of a phase that do not depend on each other (fork / join by events)
    cudaStream_t stream3 = nullptr;
#ifndef ICL_EMU
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr;
#endif
    Status st;
    int64_t n = 0;      // points
    int32_t P = 0;      // chromosome pairs
    bool uploaded = false, dp_done = false;
    icl_params prm{};
    int bits1 = 1, bits2 = 1, bitsp = 1;
    int32_t max_grad = 0;            // largest value of the uploaded grading array
    GradeTable gtab{};               // icl_set_grading_table: uploaded values are class ids, G = table[class]

    // inputs in upload order
    Buf<int32_t> in_pair, in_loc1, in_loc2, in_grad, stats, pair_hist;
    Buf<uint8_t> in_raw;             // icl_upload_typed: the narrow arrays as they crossed PCIe
    // sort
    Buf<uint64_t> key_a, key_b;
    Buf<uint32_t> idx_a, perm;
    Buf<uint8_t> tmp;
    // layout
    Buf<int64_t> d_pair_off, d_rb, d_seg_off;      // d_seg_off[2P + 1]: V-space start of every (pair, direction) item
    Buf<int32_t> d_rmin, d_nrows, d_item_order, d_pair_order;
    std::vector<int64_t> h_pair_off, h_rb, h_seg_off;
    std::vector<int32_t> h_rmin, h_nrows, h_item_order, h_pair_order;
    int64_t R = 0;      // dense rows incl. one sentinel per pair (forward); the mirrored index follows at +R
    Buf<int32_t> loc1s, loc2s, pair_s, rowptr, X1, X2, G, itemof, mmap, imap;
    // edges
    Buf<uint32_t> ecnt, eoff;
    Buf<uint2> edges;
    Buf<unsigned long long> tile_state;   // edge kernel: per-tile counts and staging places, pool cursor, per-item first records
    Buf<uint32_t> tile_dst;               // first record of every tile in the source-ordered stream
    Buf<uint2> edge_stage;                // staging pool: tiles in the order they reserved room
    uint64_t edge_cap = 0;                // records the edge pool holds
    int64_t n_edges = 0;
    int32_t max_span = 0;            // largest (target position - source position) over all edges
    int32_t chain_kernel = 0;        // 1 = shared-memory ring kernel, 0 = generic global-memory kernel
    int32_t chain_mode = 1;          // 1 = batch ring kernel (default), 3 = single-executor ring kernel, 0 = generic (env ICL_B200_CHAIN)
    Buf<long long> chain_dbg;
    int32_t *sq_fwd = nullptr, *sq_rev = nullptr, *sq_queue = nullptr;   // streaming targets of the current icl_dp_begin (host, device-visible)
    bool streamed = false, dp_begun = false;
    bool order_on_device = false;    // icl_rank_device filled order_in
    bool keys_on_device = false;     // ... from score32 (numpy flavours)
    bool order_on_device_used = false;   // the current icl_extract takes the device's order
    bool rank_flag_pending = false;      // icl_rank_device left its hand-back flag (host_flag[rank_flag_slot]) unread
    int32_t rank_flag_slot = 0;
    int32_t pvalue_mode = 0;         // 0 = by pair size, 1 = sequential warp-per-pair kernel, 2 = incidence sort (env ICL_B200_PVALUE=seq|inc)
    Buf<uint32_t> inc_cnt, inc_off;
    Buf<int32_t> inc_start;
    // DP state (V space)
    Buf<int32_t> S, root;            // root = first point of the path (root of the predecessor tree)
    Buf<int2> pd;                    // per point: {pred (-1 for path roots), path length}
    Buf<uint8_t> used;
    Buf<int32_t> score32, npscr;     // npscr: stop lists of the numpy-scalar argsort kernel
    Buf<int32_t> host_flag;          // AVX-512 argsort kernel: segments that must be ranked on the host
    Buf<uint32_t> xs_chunk;          // ... per-chunk stop counts of very long std::sort ranges
    // extraction
    Buf<int32_t> order_in, rank2v, acc_L, times, mscr;
    Buf<uint32_t> tree_v, acc_m, acc_ms;
    Buf<uint64_t> bs_in, bs;         // per emission position: accepted << 32 | wanted (m >= min_members) ... and the exclusive sums
    Buf<uint32_t> mo_in, mo;         // ... gene pairs of a wanted block ... and the exclusive sum
    Buf<int4> acc_box;               // per accepted rank position: {loc1 min, loc1 max, loc2 min, loc2 max}
    Buf<uint8_t> seen;
    Buf<uint8_t> present;
    Buf<int2> hint;                  // per point: {depth of the nearest removed ancestor, tree version} memo
    Buf<int2> big_trees;             // [t0, uend) of the trees that go to the staged walk
    Buf<int2> small_trees;           // ... and of the short ones (one thread each)
    int64_t NB = 0, NBo = 0, NMo = 0;   // accepted blocks; blocks / gene pairs the caller gets (m >= min_members)
    int32_t min_members = 1;
    Buf<int32_t> b_j0, b_j1, b_l2min, b_l2max, b_N;    // per accepted block, emission order: what the times / p-value pass needs
    Buf<int64_t> b_N1;
    // output staging (blocks the caller gets, emission order)
    Buf<uint32_t> o_src, o_ms, o_e;  // accepted-block index, member scratch slice and emission position of every output block
    icl_blocks dv{}, hv{};           // views into d_stage and into the pinned host staging (same layout)
    size_t stage_members_at = 0, stage_bytes = 0, stage_blocks_lean_end = 0, stage_members_lean_end = 0, stage_praw_at = 0, stage_late_at = 0;
    int32_t out_fields = 0;          // ICL_FIELDS_* of the current icl_extract_fields
    Buf<uint8_t> d_stage;
    uint8_t *h_stage = nullptr;
    size_t h_stage_cap = 0;
    uint64_t *h_small = nullptr;     // page-locked words for the counts a phase reads back (copies into stack variables are
                                     // staged by the driver and block the host once each)
    uint8_t *h_plan = nullptr;       // page-locked scratch of the upload: statistics down, layout up
    size_t h_plan_cap = 0;

    Timer timers[ICL_T_COUNT];
    int64_t counters[4] = {0, 0, 0, 0};
    int64_t counters_ext[2] = {0, 0};
    int64_t n_own[3] = {0, 0, 0}, n_cub[3] = {0, 0, 0};   // kernels / CUB calls of the last icl_dp, icl_rank_device, icl_extract
};

// counts the launches an ABI call makes (slot 0 icl_dp, 1 icl_rank_device, 2 icl_extract)
struct LaunchScope {
    icl_ctx *c;
    int slot;
    LaunchCount at;
    LaunchScope(icl_ctx *c_, int slot_) : c(c_), slot(slot_), at(launch_count()) { c->n_own[slot] = c->n_cub[slot] = 0; }
    ~LaunchScope() { c->n_own[slot] = launch_count().own - at.own; c->n_cub[slot] = launch_count().cub - at.cub; }
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
// input statistics (max loc1 / loc2, validity) — one warp-reduced atomic per warp — and, per chromosome pair, the
// number of points and the loc1 range: the layout of a comparison (pair offsets, dense row index, item order) is known
// when icl_upload returns, so icl_dp never has to come back to the host for it
// ------------------------------------------------------------------------------------------------
#define ICL_HIST_SMEM_PAIRS 2048     // pairs whose three counters fit a CTA's shared memory
#ifndef ICL_EMU
struct InputArrays {                 // the four uploaded arrays as they crossed PCIe (1-, 2- or 4-byte elements) and their int32 homes
    const void *raw[4];
    int32_t *dst[4];
    int32_t bytes[4];
};
__device__ __forceinline__ int32_t input_element(const InputArrays &A, int k, int64_t i) {
    const int32_t b = A.bytes[k];
    if (b == 4) return A.dst[k][i];                           // copied straight to its home
    const int32_t v = b == 1 ? (int32_t)((const uint8_t *)A.raw[k])[i] : (int32_t)((const uint16_t *)A.raw[k])[i];
    A.dst[k][i] = v;                                          // widened on the way through
    return v;
}
__global__ void __launch_bounds__(256) k_input_stats(int64_t n, int32_t P, InputArrays A, int32_t *stats, int32_t *hist) {
    // hist: [P] points, [P] smallest loc1, [P] largest loc1
    __shared__ int32_t sh[3 * ICL_HIST_SMEM_PAIRS];
    const bool staged = P <= ICL_HIST_SMEM_PAIRS;
    if (staged) {
        for (int32_t j = threadIdx.x; j < P; j += blockDim.x) { sh[j] = 0; sh[P + j] = 0x7fffffff; sh[2 * P + j] = -1; }
        __syncthreads();
    }
    int32_t *cnt = staged ? sh : hist, *lo = cnt + P, *hi = lo + P;
    int32_t m1 = 0, m2 = 0, mg = 0, bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = input_element(A, 0, i), a = input_element(A, 1, i), b = input_element(A, 2, i);
        bad |= (a < 0) | (b < 0) | (p < 0) | (p >= P);
        m1 = max(m1, a);
        m2 = max(m2, b);
        mg = max(mg, input_element(A, 3, i));
        if (p >= 0 && p < P) {
            atomicAdd(&cnt[p], 1);
            if (a < lo[p]) atomicMin(&lo[p], a);              // (the unsynchronised test only skips atomics that cannot change anything)
            if (a > hi[p]) atomicMax(&hi[p], a);
        }
    }
    if (staged) {
        __syncthreads();
        for (int32_t j = threadIdx.x; j < P; j += blockDim.x)
            if (sh[j]) { atomicAdd(&hist[j], sh[j]); atomicMin(&hist[P + j], sh[P + j]); atomicMax(&hist[2 * P + j], sh[2 * P + j]); }
    }
    m1 = __reduce_max_sync(0xffffffffu, m1);
    m2 = __reduce_max_sync(0xffffffffu, m2);
    mg = __reduce_max_sync(0xffffffffu, mg);
    bad = __reduce_or_sync(0xffffffffu, (unsigned)bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&stats[0], m1);
        atomicMax(&stats[1], m2);
        atomicMax(&stats[3], mg);
        if (bad) atomicOr(&stats[2], 1);
    }
}
#endif

static int input_stats(icl_ctx *c, const uint8_t *const raw[4], const int32_t bytes[4]) {
    Status &st = c->st;
    const int32_t P = c->P;
    if (c->pair_hist.reserve(st, 3 * (size_t)P + 1)) return st.code;
    ICL_CUDA(st, dev_memset(c->stats.p, 0, 4 * sizeof(int32_t), c->stream));
    {
        int32_t *hist = c->pair_hist.p;
        ICL_CUDA(st, foreach(c->stream, 3 * (int64_t)P, ICL_LAMBDA(int64_t j) { hist[j] = j < P ? 0 : j < 2 * (int64_t)P ? 0x7fffffff : -1; }));
    }
#ifndef ICL_EMU
    if (c->n > 0) {
        int64_t blocks = std::min<int64_t>((c->n + 255) / 256, 148 * 8);
        InputArrays A;
        int32_t *dst[4] = {c->in_pair.p, c->in_loc1.p, c->in_loc2.p, c->in_grad.p};
        for (int k = 0; k < 4; k++) { A.raw[k] = raw[k]; A.dst[k] = dst[k]; A.bytes[k] = bytes[k]; }
        k_input_stats<<<(unsigned)blocks, 256, 0, c->stream>>>(c->n, P, A, c->stats.p, c->pair_hist.p);
        ICL_COUNT_KERNEL();
        ICL_CUDA(st, cudaGetLastError());
    }
#else
    {
        int32_t *dst[4] = {c->in_pair.p, c->in_loc1.p, c->in_loc2.p, c->in_grad.p};
        for (int k = 0; k < 4; k++)
            for (int64_t i = 0; i < c->n && bytes[k] != 4; i++)
                dst[k][i] = bytes[k] == 1 ? (int32_t)raw[k][i] : (int32_t)((const uint16_t *)raw[k])[i];
    }
    for (int64_t i = 0; i < c->n; i++) {
        int32_t a = c->in_loc1.p[i], b = c->in_loc2.p[i], p = c->in_pair.p[i];
        if (a < 0 || b < 0 || p < 0 || p >= P) c->stats.p[2] = 1;
        c->stats.p[0] = std::max(c->stats.p[0], a);
        c->stats.p[1] = std::max(c->stats.p[1], b);
        c->stats.p[3] = std::max(c->stats.p[3], c->in_grad.p[i]);
        if (p >= 0 && p < P) {
            int32_t *h = c->pair_hist.p;
            h[p]++;
            h[P + p] = std::min(h[P + p], a);
            h[2 * P + p] = std::max(h[2 * P + p], a);
        }
    }
#endif
    return 0;
}

// page-locked host scratch of the upload (grow-only)
static int host_scratch(icl_ctx *c, size_t bytes) {
    Status &st = c->st;
    if (bytes <= c->h_plan_cap) return 0;
    if (c->h_plan) pin_free(c->h_plan);
    c->h_plan = nullptr; c->h_plan_cap = 0;
    if (pin_malloc((void **)&c->h_plan, 2 * bytes) != cudaSuccess || !c->h_plan) {
        c->h_plan = nullptr;
        return fail(c, ICL_ERR_NOMEM, "icl_upload: out of page-locked host memory");
    }
    c->h_plan_cap = 2 * bytes;
    return 0;
}

// The layout of the uploaded comparison, from the per-pair counts of input_stats: where every pair's points will stand
// after the sort, its dense row range, the longest-processing-time order of pairs / (pair, direction) items.
static int plan_layout(icl_ctx *c, const int32_t *hist, const char *who) {
    Status &st = c->st;
    const int64_t n = c->n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    c->h_pair_off.assign(P + 1, 0);
    c->h_rmin.assign(P + 1, 0);
    c->h_nrows.assign(P + 1, 0);
    for (int32_t p = 0; p < P; p++) {
        c->h_pair_off[p + 1] = c->h_pair_off[p] + hist[p];
        if (hist[p] > 0) { c->h_rmin[p] = hist[P + p]; c->h_nrows[p] = hist[2 * P + p] - hist[P + p] + 1; }
    }
    // dense row index: nrows[p] + 1 slots per pair
    c->h_rb.assign(P + 1, 0);
    for (int32_t p = 0; p < P; p++) c->h_rb[p + 1] = c->h_rb[p] + c->h_nrows[p] + 1;
    c->R = c->h_rb[P];
    for (int32_t p = 0; p < P; p++)
        if (hist[p] >= (1 << 24)) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": a chromosome pair must hold fewer than 2^24 points");
    if (2 * c->R >= (1ll << 31)) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": dense row index too large (sparse gene orders?)");
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
    c->h_seg_off.resize(2 * (size_t)P + 1);
    for (int32_t p = 0; p <= P; p++) { c->h_seg_off[p] = po[p]; c->h_seg_off[P + p] = n + po[p]; }
    if (c->d_pair_off.reserve(st, P + 1) || c->d_rmin.reserve(st, P + 1) || c->d_nrows.reserve(st, P + 1) || c->d_rb.reserve(st, P + 1) ||
        c->d_item_order.reserve(st, 2 * P + 1) || c->d_pair_order.reserve(st, P + 1) || c->d_seg_off.reserve(st, 2 * P + 1))
        return st.code;
    // the copies leave from the page-locked scratch (behind the counters just read): asynchronous, nothing to wait for —
    // the scratch is next written by the next upload, which starts with a synchronisation of its own
    uint8_t *at = c->h_plan + (((4 + 3 * (size_t)P) * sizeof(int32_t) + 15) & ~(size_t)15);
    auto send = [&](void *dst, const void *src, size_t bytes) {
        memcpy(at, src, bytes);
        cudaError_t e = h2d(dst, at, bytes, s);
        at += (bytes + 15) & ~(size_t)15;
        return e;
    };
    ICL_CUDA(st, send(c->d_pair_off.p, c->h_pair_off.data(), (size_t)(P + 1) * sizeof(int64_t)));
    ICL_CUDA(st, send(c->d_seg_off.p, c->h_seg_off.data(), (2 * (size_t)P + 1) * sizeof(int64_t)));
    ICL_CUDA(st, send(c->d_rb.p, c->h_rb.data(), (size_t)(P + 1) * sizeof(int64_t)));
    ICL_CUDA(st, send(c->d_rmin.p, c->h_rmin.data(), (size_t)(P + 1) * sizeof(int32_t)));
    ICL_CUDA(st, send(c->d_nrows.p, c->h_nrows.data(), (size_t)(P + 1) * sizeof(int32_t)));
    if (P > 0) {
        ICL_CUDA(st, send(c->d_item_order.p, c->h_item_order.data(), (size_t)2 * P * sizeof(int32_t)));
        ICL_CUDA(st, send(c->d_pair_order.p, c->h_pair_order.data(), (size_t)P * sizeof(int32_t)));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: lifetime
// ------------------------------------------------------------------------------------------------
extern "C" void *icl_stream(icl_ctx *c) { return c ? (void *)c->stream : nullptr; }

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
    e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete c; return ICL_ERR_CUDA; }
#endif
    if (pin_malloc((void **)&c->h_small, 64 * sizeof(uint64_t)) != cudaSuccess) c->h_small = nullptr;
    if (!c->h_small) { g_create_error = "icl_create: no page-locked host memory"; delete c; return ICL_ERR_CUDA; }
    if (c->stats.reserve(c->st, 8)) { g_create_error = c->st.msg; delete c; return ICL_ERR_CUDA; }
    const char *mode = getenv("ICL_B200_CHAIN");              // diagnostics: A/B the DP kernels
    const char *pvm = getenv("ICL_B200_PVALUE");
    if (pvm) c->pvalue_mode = !strcmp(pvm, "seq") ? 1 : !strcmp(pvm, "inc") ? 2 : !strcmp(pvm, "warp") ? 3 : 0;
    if (mode) c->chain_mode = !strcmp(mode, "generic") ? 0 : !strcmp(mode, "single") ? 3 : 1;
    *out = c;
    return ICL_OK;
}

extern "C" void icl_destroy(icl_ctx *c) {
    if (!c) return;
#ifndef ICL_EMU
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream2);           // released blocks go to the process-wide cache (icl_platform.h): nothing of
    cudaStreamSynchronize(c->stream3);           // this context may still be running when another context picks them up
#endif
    c->in_pair.release(); c->in_loc1.release(); c->in_loc2.release(); c->in_grad.release(); c->stats.release(); c->in_raw.release();
    c->key_a.release(); c->key_b.release(); c->idx_a.release(); c->perm.release(); c->tmp.release();
    c->d_pair_off.release(); c->d_rb.release(); c->d_seg_off.release(); c->d_rmin.release(); c->d_nrows.release();
    c->d_item_order.release(); c->d_pair_order.release();
    c->loc1s.release(); c->loc2s.release(); c->pair_s.release(); c->rowptr.release();
    c->X1.release(); c->X2.release(); c->G.release(); c->itemof.release(); c->mmap.release(); c->imap.release();
    c->ecnt.release(); c->eoff.release(); c->edges.release(); c->tile_state.release(); c->tile_dst.release(); c->edge_stage.release(); c->chain_dbg.release(); c->inc_cnt.release(); c->inc_off.release(); c->inc_start.release();
    c->S.release(); c->root.release(); c->hint.release(); c->big_trees.release(); c->small_trees.release(); c->pd.release(); c->used.release(); c->score32.release();
    c->order_in.release(); c->mscr.release(); c->acc_ms.release(); c->acc_box.release(); c->seen.release(); c->rank2v.release(); c->acc_L.release();
    c->times.release(); c->tree_v.release(); c->acc_m.release(); c->bs_in.release(); c->bs.release(); c->mo_in.release(); c->mo.release();
    c->present.release(); c->b_j0.release();
    c->b_j1.release(); c->b_l2min.release(); c->b_l2max.release();
    c->b_N.release(); c->b_N1.release();
    c->d_stage.release();
    c->o_src.release(); c->o_ms.release(); c->o_e.release();
    c->host_flag.release(); c->xs_chunk.release(); c->npscr.release();
    if (c->h_stage) pin_free(c->h_stage);
    if (c->h_plan) pin_free(c->h_plan);
    if (c->h_small) pin_free(c->h_small);
    c->pair_hist.release();
    for (auto &t : c->timers) t.destroy();
#ifndef ICL_EMU
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream3);
    cudaEventDestroy(c->ev_join3);
    cudaStreamDestroy(c->stream3);
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->stream2);
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
#endif
    *ptr = nullptr;
    if (pin_malloc(ptr, (size_t)(bytes ? bytes : 16)) != cudaSuccess || !*ptr) return fail(c, ICL_ERR_NOMEM, "icl_host_alloc: out of memory");
    return ICL_OK;
}

extern "C" int icl_host_free(icl_ctx *c, void *ptr) {
    if (!c) return ICL_ERR_ARG;
#ifndef ICL_EMU
    ICL_CUDA(c->st, cudaSetDevice(c->device));
#endif
    pin_free(ptr);
    return ICL_OK;
}

// ------------------------------------------------------------------------------------------------
// ABI: upload
// ------------------------------------------------------------------------------------------------
// the four input arrays in 1-, 2- or 4-byte unsigned elements: copied as they are, widened to int32 on the device
static int upload_typed(icl_ctx *c, int64_t n, int32_t P, const void *const src[4], const int32_t bytes[4], const char *who) {
    c->uploaded = c->dp_done = c->dp_begun = false;
    if (n < 0 || P < 0 || (n > 0 && (!src[0] || !src[1] || !src[2] || !src[3])) || (n > 0 && P == 0))
        return fail(c, ICL_ERR_ARG, std::string(who) + ": bad argument");
    for (int k = 0; k < 4; k++)
        if (bytes[k] != 1 && bytes[k] != 2 && bytes[k] != 4) return fail(c, ICL_ERR_ARG, std::string(who) + ": element sizes must be 1, 2 or 4 bytes");
    if (n >= (1ll << 30)) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": n_points must be < 2^30");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    c->n = n;
    c->P = P;
    if (c->in_pair.reserve(st, n) || c->in_loc1.reserve(st, n) || c->in_loc2.reserve(st, n) || c->in_grad.reserve(st, n))
        return st.code;
    int32_t *dst[4] = {c->in_pair.p, c->in_loc1.p, c->in_loc2.p, c->in_grad.p};
    size_t narrow = 0;
    for (int k = 0; k < 4; k++) if (bytes[k] != 4) narrow += ((size_t)n * bytes[k] + 15) & ~(size_t)15;
    if (narrow && c->in_raw.reserve(st, narrow)) return st.code;
    uint8_t *raw = c->in_raw.p;
    const uint8_t *raws[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; k++) {
        if (bytes[k] == 4) { ICL_CUDA(st, h2d(dst[k], src[k], (size_t)n * 4, c->stream)); continue; }
        ICL_CUDA(st, h2d(raw, src[k], (size_t)n * bytes[k], c->stream));
        raws[k] = raw;                                         // widened by the statistics pass
        raw += ((size_t)n * bytes[k] + 15) & ~(size_t)15;
    }
    // page-locked scratch for the statistics coming back and the layout going up: [4] stats, [3P] per-pair counters, plan
    const size_t plan_bytes = (4 + 3 * (size_t)P) * sizeof(int32_t) + (4 * (size_t)P + 8) * sizeof(int64_t) + (5 * (size_t)P + 8) * sizeof(int32_t) + 256;
    if (host_scratch(c, plan_bytes)) return st.code;
    int32_t *hs = (int32_t *)c->h_plan, *hist = hs + 4;
    if (input_stats(c, raws, bytes)) return st.code;
    ICL_CUDA(st, d2h(hs, c->stats.p, 4 * sizeof(int32_t), c->stream));
    if (P > 0) ICL_CUDA(st, d2h(hist, c->pair_hist.p, 3 * (size_t)P * sizeof(int32_t), c->stream));
    ICL_CUDA(st, dev_sync(c->stream));
    if (hs[2]) return fail(c, ICL_ERR_ARG, std::string(who) + ": negative coordinate or pair id outside [0, n_pairs)");
    if (hs[0] >= (1 << 26) || hs[1] >= (1 << 26)) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": coordinates must be < 2^26");
    if (hs[3] > 255) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": grading values must be <= 255");
    c->max_grad = hs[3];
    c->bits1 = bits_for(hs[0]);
    c->bits2 = bits_for(hs[1]);
    c->bitsp = bits_for(P > 0 ? P - 1 : 0);
    if (c->bits1 + c->bits2 + c->bitsp > 64) return fail(c, ICL_ERR_LIMIT, std::string(who) + ": key does not fit 64 bits");
    if (plan_layout(c, hist, who)) return st.code;
    c->uploaded = true;
    st.code = 0;
    return ICL_OK;
}

extern "C" int icl_upload(icl_ctx *c, int64_t n, int32_t P, const int32_t *pair_id, const int32_t *loc1,
                          const int32_t *loc2, const int32_t *grading) {
    if (!c) return ICL_ERR_ARG;
    const void *src[4] = {pair_id, loc1, loc2, grading};
    const int32_t bytes[4] = {4, 4, 4, 4};
    return upload_typed(c, n, P, src, bytes, "icl_upload");
}

extern "C" int icl_upload_typed(icl_ctx *c, int64_t n, int32_t P, const void *pair_id, int32_t pair_bytes, const void *loc1,
                                int32_t loc1_bytes, const void *loc2, int32_t loc2_bytes, const void *grading, int32_t grading_bytes) {
    if (!c) return ICL_ERR_ARG;
    const void *src[4] = {pair_id, loc1, loc2, grading};
    const int32_t bytes[4] = {pair_bytes, loc1_bytes, loc2_bytes, grading_bytes};
    return upload_typed(c, n, P, src, bytes, "icl_upload_typed");
}

// ------------------------------------------------------------------------------------------------
// ABI: grading classes.  The grading of a hit is a function of its per-query rank class only
// (wgdi/run_colliearity.py:64-86: red / blue / rest), so a parameter sweep over `grading` re-maps the classes of the
// SAME upload instead of uploading again.
// ------------------------------------------------------------------------------------------------
extern "C" int icl_set_grading_table(icl_ctx *c, int32_t n_classes, const int32_t *values) {
    if (!c) return ICL_ERR_ARG;
    if (n_classes < 0 || n_classes > ICL_MAX_GRADE_CLASSES || (n_classes > 0 && !values))
        return fail(c, ICL_ERR_ARG, "icl_set_grading_table: bad argument");
    for (int32_t k = 0; k < n_classes; k++)
        if (values[k] < 0 || values[k] > 255) return fail(c, ICL_ERR_LIMIT, "icl_set_grading_table: grading values must be in 0..255");
    c->gtab.n = n_classes;
    for (int32_t k = 0; k < ICL_MAX_GRADE_CLASSES; k++) c->gtab.v[k] = k < n_classes ? values[k] : 0;
    return ICL_OK;
}

// ------------------------------------------------------------------------------------------------
// phase: sort + layout
// ------------------------------------------------------------------------------------------------
struct SortedKeys {                  // the sorted (pair, loc1, loc2) keys, 64- or 32-bit
    const void *p;
    bool wide;
    ICL_HD uint64_t operator[](int64_t i) const { return wide ? ((const uint64_t *)p)[i] : (uint64_t)((const uint32_t *)p)[i]; }
};

static int phase_sort(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n;
    cudaStream_t s = c->stream;
    if (c->key_a.reserve(st, n) || c->key_b.reserve(st, n) || c->idx_a.reserve(st, n) || c->perm.reserve(st, n) ||
        c->loc1s.reserve(st, n) || c->loc2s.reserve(st, n) || c->pair_s.reserve(st, n) || c->G.reserve(st, 2 * n))
        return st.code;
    const int sh2 = c->bits2, shp = c->bits1 + c->bits2;
    // (pair, loc1, loc2) of a Vitis- or wheat-sized comparison needs 31-32 bits: 32-bit keys then (a third less traffic per pass)
    const bool wide = shp + c->bitsp > 32;
    {
        const int32_t *ip = c->in_pair.p, *i1 = c->in_loc1.p, *i2 = c->in_loc2.p;
        uint64_t *key = c->key_a.p;
        uint32_t *key32 = (uint32_t *)c->key_a.p;
        uint32_t *idx = c->idx_a.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            const uint64_t k = ((uint64_t)(uint32_t)ip[i] << shp) | ((uint64_t)(uint32_t)i1[i] << sh2) | (uint64_t)(uint32_t)i2[i];
            if (wide) key[i] = k; else key32[i] = (uint32_t)k;
            idx[i] = (uint32_t)i;
        }));
    }
    // stable LSD radix sort: equal (pair, loc1, loc2) keep upload order (run_colliearity.py:164-166)
    if (n > 0 && wide && radix_sort_pairs(st, c->tmp, c->key_a.p, c->key_b.p, c->idx_a.p, c->perm.p, n, shp + c->bitsp, s)) return st.code;
    if (n > 0 && !wide && radix_sort_pairs32(st, c->tmp, (const uint32_t *)c->key_a.p, (uint32_t *)c->key_b.p, c->idx_a.p, c->perm.p, n,
                                             shp + c->bitsp, s))
        return st.code;
    {
        const SortedKeys key{c->key_b.p, wide};
        const uint32_t *perm = c->perm.p;
        const int32_t *ig = c->in_grad.p;
        int32_t *l1 = c->loc1s.p, *l2 = c->loc2s.p, *ps = c->pair_s.p, *G = c->G.p;
        const uint64_t m1 = (1ull << c->bits1) - 1, m2 = (1ull << c->bits2) - 1;
        const GradeTable gt = c->gtab;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            uint64_t k = key[i];
            l2[i] = (int32_t)(k & m2);
            l1[i] = (int32_t)((k >> sh2) & m1);
            ps[i] = (int32_t)(k >> shp);
            const int32_t g = ig[perm[i]];
            G[i] = gt.n ? gt.v[g] : g;                        // class id -> grading of this parameter set
        }));
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
    // band of source v: the points of the next mg1-1 rows, one contiguous slice [k0, k1) of V space
    ICL_HD bool band(int64_t v, int32_t &k0, int32_t &k1, int32_t &x1, int32_t &x2, int64_t &base, int32_t &nr) const {
        int32_t it = itemof[v];
        int32_t p = it >= P ? it - P : it;
        base = rb[p] + (it >= P ? R : 0);
        nr = nrows[p];
        x1 = X1[v];
        x2 = X2[v];
        int32_t lastrow = x1 + mg1 - 1 < nr - 1 ? x1 + mg1 - 1 : nr - 1;
        if (lastrow <= x1) { k0 = k1 = 0; return false; }
        k0 = rowptr[base + x1 + 1];
        k1 = rowptr[base + lastrow + 1];
        return k1 > k0;
    }
    // candidates of source v in visiting order; pX1/pX2/pG are indexed by V position (global arrays or a staged tile)
    template <class Emit>
    ICL_HD void scan(int32_t k0, int32_t k1, int32_t x1, int32_t x2, int64_t base, int32_t nr,
                     const int32_t *pX1, const int32_t *pX2, const int32_t *pG, Emit emit) const {
        if (k1 - k0 <= 384) {
            // the next mg1-1 rows are one contiguous slice sorted by (row, column): scan it once
            for (int32_t k = k0; k < k1; k++) {
                int32_t dc = pX2[k] - x2;
                if (dc <= 0 || dc >= mg2) continue;
                int32_t dr = pX1[k] - x1;
                int32_t s = pG[k] - (dr + dc);
                bool last = (k + 1 == k1) || pX1[k + 1] != pX1[k] || (pX2[k + 1] - x2 >= mg2);
                if (s > 0 || last) emit(k, dr, dc, s > 0 ? s : 0);
            }
            return;
        }
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
    template <class Emit>
    ICL_HD void run(int64_t v, Emit emit) const {
        int32_t k0, k1, x1, x2, nr;
        int64_t base;
        if (!band(v, k0, k1, x1, x2, base, nr)) return;
        scan(k0, k1, x1, x2, base, nr, X1, X2, G, emit);
    }
};

// Edge record (8 bytes) of the flat per-item edge stream the DP kernel consumes:
//   x = koff (bits 0-15: target position - source position) | dr << 16 | dc << 24
//   y = s (bits 0-7: grading_k - (dr + dc), 0 when it cannot update) | srel << 8
// srel (24 bits) = source position relative to the first point of the item, so a record is self-contained: a
// warp can replay 32 consecutive records without first scanning for their sources.  Sources without
// candidates have no records and cost the DP kernel nothing.
#define ICL_EDGE_CHUNK 512           // records per TMA chunk (4 KB)

#ifndef ICL_EMU
// Edge enumeration: 256 consecutive sources per CTA (one thread each); every record is built ONCE.
//  * Consecutive sources have heavily overlapping bands (the points are sorted), so the CTA stages the union of its
//    bands — a contiguous slice of X1 / X2 / G — in shared memory.
//  * stage 1: every thread scans ITS band for the column window with 16-byte shared-memory loads and does nothing on a
//    hit but set a bit (a predicated OR): the scan stays convergent.  (ncu on the two-pass predecessor: issue-bound,
//    431 M warp instructions, 10-16 of 32 threads active around every hit; a hit LIST appended under a branch still spent
//    28 % of its instructions in the append at 11.6 threads.)
//  * stage 2: the thread walks the set bits of a word (a few) and keeps the candidates the sweep can see (s > 0 or
//    row-last, see above); a block scan numbers them inside the tile.
//  * stage 3: the tile reserves room in a staging pool with one atomic and its threads store their records there,
//    contiguous and in source order INSIDE the tile.  Tiles land in the pool in any order; k_edges_place then moves every
//    tile's run to its place in the source-ordered stream (a coalesced copy of 8 E bytes; the tile offsets are one small
//    scan).  A chained scan with look-back inside this kernel was tried first: correct, but every tile waited for the
//    slowest tile before it (barrier stall 12.8 warps per issue, 48 % of the samples in the look-back: profiles/).
// Only the record index at the first point of every (pair, direction) item is kept (eoff[item start], eoff[V] = total):
// records name their source, the DP kernel needs nothing else.
#define ICL_EDGE_TILE 2048
#define ICL_EDGE_BAND 384            // bands up to this many points are scanned from the staged slice
#define ICL_EDGE_WORDS ((ICL_EDGE_BAND + 3 + 31) / 32 + 1)   // candidate-bit words per thread
struct EdgeScan {
    uint32_t *tile_cnt;              // [tiles + 1] records of the tile
    uint32_t *tile_at;               // [tiles] where the tile's run starts in the staging pool
    uint32_t *first_local;           // [2P] tile-local record index of the first source of every item
    unsigned long long *cursor;      // staging pool fill (may exceed cap: then the caller grows the pool and repeats)
    uint2 *staging;
    uint64_t cap;                    // records the pool holds
    const int64_t *seg_off;          // [2P + 1] first V position of every item
};

__global__ void __launch_bounds__(256, 4) k_edges_enum(EdgeEnum ee, int64_t V, EdgeScan es, int32_t *stats) {
    __shared__ __align__(16) int32_t sX2[ICL_EDGE_TILE + 40];
    __shared__ int32_t sX1[ICL_EDGE_TILE + 8], sG[ICL_EDGE_TILE + 8];
    __shared__ uint32_t hm[ICL_EDGE_WORDS * 256];             // per thread: bit e of word w = band position a4 + 32 w + e is a candidate
    __shared__ int32_t s_kmin, s_kmax, s_span;
    __shared__ unsigned long long s_base;
    typedef cub::BlockScan<uint32_t, 256> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    const int tid = threadIdx.x;
    if (tid == 0) { s_kmin = 0x7fffffff; s_kmax = 0; s_span = 0; }
    __syncthreads();
    const uint32_t tile = blockIdx.x;
    const int64_t v = (int64_t)tile * 256 + tid;
    int32_t k0 = 0, k1 = 0, x1 = 0, x2 = 0, nr = 0;
    int64_t base = 0;
    bool has = false;
    if (v < V) has = ee.band(v, k0, k1, x1, x2, base, nr);
    const bool small = has && (k1 - k0 <= ICL_EDGE_BAND);
    {
        int32_t lo = small ? k0 : 0x7fffffff, hi = small ? k1 : 0;
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((tid & 31) == 0 && hi > lo) { atomicMin(&s_kmin, lo); atomicMax(&s_kmax, hi); }
    }
    __syncthreads();
    const int32_t kbase = s_kmin & ~3, kmax = s_kmax;          // 16-byte alignment of the staged slice follows the global one
    const bool tiled = kmax > kbase && kmax - kbase <= ICL_EDGE_TILE;
    if (tiled) {
        for (int32_t k = kbase + tid; k < kmax; k += 256) { sX1[k - kbase] = ee.X1[k]; sX2[k - kbase] = ee.X2[k]; sG[k - kbase] = ee.G[k]; }
    }
    __syncthreads();
    const int32_t mg2 = ee.mg2;
    const bool fast = tiled && small;
    const int32_t a = k0 - kbase, b = k1 - kbase;              // the band inside the staged slice (fast only)
    const int32_t a4 = a & ~3;
    const int32_t nwords = fast ? (b - a4 + 31) >> 5 : 0;
    uint32_t cnt = 0;                                          // records of this source
    int32_t span = 0;
    if (fast) {
        // ---- stage 1: which positions of the band have their column in (x2, x2 + mg2): one bit each, no branch on a hit
        const uint32_t w = (uint32_t)(mg2 - 1);
        const int32_t xs = x2 + 1;
        for (int32_t wd = 0; wd < nwords; wd++) {
            const int32_t j0 = a4 + (wd << 5);
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (j0 + 4 * u >= b) break;
                const int4 q = *(const int4 *)(sX2 + j0 + 4 * u);
                if ((uint32_t)(q.x - xs) < w) m |= 1u << (4 * u);
                if ((uint32_t)(q.y - xs) < w) m |= 2u << (4 * u);
                if ((uint32_t)(q.z - xs) < w) m |= 4u << (4 * u);
                if ((uint32_t)(q.w - xs) < w) m |= 8u << (4 * u);
            }
            if (wd == 0) m &= ~0u << (a - a4);                 // positions before the band
            const int32_t rem = b - j0;
            if (rem < 32) m &= (1u << rem) - 1u;               // positions behind it
            // ---- stage 2: keep what the sweep can see — s > 0, or the last candidate of its row — in visiting order
            uint32_t t = m;
            while (t) {
                const int e = __ffs(t) - 1;
                t &= t - 1;
                const int32_t q = j0 + e, r = sX1[q];
                const int32_t sc = sG[q] - (r - x1 + sX2[q] - x2);
                if (sc > 0 || (q + 1 == b) || sX1[q + 1] != r || (sX2[q + 1] - x2 >= mg2)) span = kbase + q - (int32_t)v;
                else m &= ~(1u << e);
            }
            hm[wd * 256 + tid] = m;
            cnt += __popc(m);
        }
    } else if (has) {                                          // oversized band or tile: the generic scan on global memory
        ee.scan(k0, k1, x1, x2, base, nr, ee.X1, ee.X2, ee.G, [&](int32_t k, int32_t, int32_t, int32_t) { cnt++; span = k - (int32_t)v; });
    }
    uint32_t off_in_tile, tile_total;
    Scan(scan_tmp).ExclusiveSum(cnt, off_in_tile, tile_total);
    if (span > 0) atomicMax(&s_span, span);
    if (tid == 0) {
        const unsigned long long at = tile_total ? atomicAdd(es.cursor, (unsigned long long)tile_total) : 0ull;
        s_base = at;
        es.tile_cnt[tile] = tile_total;
        es.tile_at[tile] = (uint32_t)at;
    }
    __syncthreads();
    if (tid == 0 && s_span > stats[3]) atomicMax(&stats[3], s_span);
    if (v >= V) return;
    const int32_t it = ee.itemof[v];
    const int64_t istart = es.seg_off[it];
    if (v == istart) es.first_local[it] = off_in_tile;
    if (cnt == 0) return;
    const unsigned long long e0 = s_base + off_in_tile;
    if (e0 + cnt > es.cap) return;                             // pool too small: the caller sees the cursor and repeats
    const uint32_t srel = (uint32_t)(v - istart) << 8;
    uint2 *out = es.staging + e0;
    // ---- stage 3: the records of the kept candidates
    if (fast) {
        for (int32_t wd = 0; wd < nwords; wd++) {
            uint32_t t = hm[wd * 256 + tid];
            while (t) {
                const int32_t q = a4 + (wd << 5) + __ffs(t) - 1;
                t &= t - 1;
                const int32_t dr = sX1[q] - x1, dc = sX2[q] - x2;
                const int32_t sc = sG[q] - (dr + dc);
                *out++ = make_uint2(((uint32_t)(kbase + q - (int32_t)v) & 0xffffu) | ((uint32_t)dr << 16) | ((uint32_t)dc << 24),
                                    ((uint32_t)(sc > 0 ? sc : 0) & 0xffu) | srel);
            }
        }
    } else {
        ee.scan(k0, k1, x1, x2, base, nr, ee.X1, ee.X2, ee.G, [&](int32_t k, int32_t dr, int32_t dc, int32_t sc) {
            *out++ = make_uint2(((uint32_t)(k - (int32_t)v) & 0xffffu) | ((uint32_t)dr << 16) | ((uint32_t)dc << 24),
                                ((uint32_t)sc & 0xffu) | srel);
        });
    }
}

// every tile's run of records from where it landed in the staging pool to its place in the source-ordered stream
__global__ void __launch_bounds__(256) k_edges_place(uint32_t tiles, const uint32_t *tile_cnt, const uint32_t *tile_at,
                                                     const uint32_t *tile_dst, const uint2 *staging, uint2 *edges,
                                                     const unsigned long long *cursor, uint64_t cap) {
    if (*cursor > cap) return;                                 // the staging pool overflowed: the caller repeats the enumeration
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t n = tile_cnt[t];
        const uint2 *src = staging + tile_at[t];
        uint2 *dst = edges + tile_dst[t];
        for (uint32_t i = threadIdx.x; i < n; i += 256) dst[i] = src[i];
    }
}
#endif

static int phase_edges(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->eoff.reserve(st, V + 2)) return st.code;
    EdgeEnum ee{c->X1.p, c->X2.p, c->G.p, c->itemof.p, c->rowptr.p, c->d_nrows.p, c->d_rb.p, c->R, c->P, c->prm.mg1, c->prm.mg2};
    int32_t *stats = c->stats.p;
    ICL_CUDA(st, dev_memset(c->stats.p + 3, 0, sizeof(int32_t), s));
#ifndef ICL_EMU
    const unsigned tiles = (unsigned)((V + 255) / 256);
    uint64_t total = 0;
    int32_t span = 0;
    if (V > 0) {
        // per tile: count, place in the staging pool, place in the stream; then cursor (8 B, aligned) and per-item first records
        if (c->tile_state.reserve(st, 2 * (size_t)tiles + (size_t)P + 8) || c->tile_dst.reserve(st, (size_t)tiles + 2)) return st.code;
        uint32_t *tile_cnt = (uint32_t *)c->tile_state.p, *tile_at = tile_cnt + tiles + 2;
        unsigned long long *cursor = c->tile_state.p + tiles + 2;
        uint32_t *first_local = (uint32_t *)(cursor + 1);
        // the pools are sized from experience (records per source: C2 1.3, wide windows a few times that) and grown when
        // a comparison needs more: the kernel reports the real total and stores nothing that does not fit
        uint64_t want = std::max<uint64_t>(c->edge_cap, 3 * (uint64_t)V + 4096);
        for (int attempt = 0; attempt < 2; attempt++) {
            if (c->edges.reserve(st, (size_t)want + 2 * ICL_EDGE_CHUNK + 16) || c->edge_stage.reserve(st, (size_t)want + 16)) return st.code;
            c->edge_cap = want;
            ICL_CUDA(st, dev_memset(cursor, 0, sizeof(unsigned long long), s));
            EdgeScan es{tile_cnt, tile_at, first_local, cursor, c->edge_stage.p, want, c->d_seg_off.p};
            k_edges_enum<<<tiles, 256, 0, s>>>(ee, V, es, stats);
            ICL_COUNT_KERNEL();
            ICL_CUDA(st, cudaGetLastError());
            // total and window span are final here: they travel to the host on the side stream while the tile scan and the
            // placement run, so the host is awake and the DP kernel queued by the time those finish
            ICL_CUDA(st, cudaEventRecord(c->ev_fork, s));
            ICL_CUDA(st, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
            ICL_CUDA(st, d2h(&c->h_small[0], cursor, sizeof(uint64_t), c->stream2));
            ICL_CUDA(st, d2h(&c->h_small[1], c->stats.p + 3, sizeof(int32_t), c->stream2));
            ICL_CUDA(st, cudaEventRecord(c->ev_join, c->stream2));
            if (exclusive_sum_u32(st, c->tmp, tile_cnt, c->tile_dst.p, tiles, s)) return st.code;
            const uint32_t *tile_dst = c->tile_dst.p;
            // (a pool that turned out too small holds nothing beyond its end: the placement is skipped on the device and the
            // whole enumeration repeated below)
            k_edges_place<<<std::min<unsigned>(tiles, 148 * 8), 256, 0, s>>>(tiles, tile_cnt, tile_at, tile_dst, c->edge_stage.p, c->edges.p, cursor, want);
            ICL_COUNT_KERNEL();
            ICL_CUDA(st, cudaGetLastError());
            {
                // record index at the first point of every item (the DP kernel reads eoff[item start] and eoff[next start])
                uint32_t *eoff = c->eoff.p;
                const int64_t *seg_off = c->d_seg_off.p;
                ICL_CUDA(st, foreach(s, 2 * (int64_t)P + 1, ICL_LAMBDA(int64_t it) {
                    if (it == 2 * (int64_t)P) { eoff[V] = tile_dst[tiles]; return; }
                    const int64_t a = seg_off[it];
                    if (seg_off[it + 1] > a) eoff[a] = tile_dst[a >> 8] + first_local[it];
                }));
            }
            ICL_CUDA(st, cudaEventSynchronize(c->ev_join));
            total = c->h_small[0];
            span = *(const int32_t *)&c->h_small[1];
            if (total <= want) break;
            want = total + total / 8 + 4096;                   // enumerate again into pools that hold everything
        }
        if (total >= (1ull << 32)) return fail(c, ICL_ERR_LIMIT, "icl_dp: more than 2^32 candidate records");
    }
    c->n_edges = (int64_t)total;
    c->max_span = span;
    if (span >= 65536) return fail(c, ICL_ERR_LIMIT, "icl_dp: more than 65535 points inside one mg1-row window of a chromosome pair");
#else
    if (c->ecnt.reserve(st, V + 2)) return st.code;
    uint32_t *ecnt = c->ecnt.p;
    ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) {
        uint32_t cnt = 0;
        int32_t span = 0;
        ee.run(v, [&](int32_t k, int32_t, int32_t, int32_t) { cnt++; span = k - (int32_t)v; });
        ecnt[v] = cnt;
        if (span > stats[3]) atomic_max_i32(&stats[3], span);
    }));
    if (exclusive_sum_u32(st, c->tmp, c->ecnt.p, c->eoff.p, V, s)) return st.code;
    uint32_t total = 0;
    int32_t span = 0;
    ICL_CUDA(st, d2h(&total, c->eoff.p + V, sizeof(uint32_t), s));
    ICL_CUDA(st, d2h(&span, c->stats.p + 3, sizeof(int32_t), s));
    ICL_CUDA(st, dev_sync(s));
    c->n_edges = total;
    c->max_span = span;
    if (span >= 65536) return fail(c, ICL_ERR_LIMIT, "icl_dp: more than 65535 points inside one mg1-row window of a chromosome pair");
    if (c->edges.reserve(st, (size_t)total + 2 * ICL_EDGE_CHUNK + 16)) return st.code;
    const uint32_t *eoff = c->eoff.p;
    uint2 *edges = c->edges.p;
    const int32_t *itemof = c->itemof.p;
    const int64_t *off = c->d_pair_off.p;
    ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) {
        uint32_t e = eoff[v];
        if (eoff[v + 1] == e) return;
        int32_t it = itemof[v];
        int64_t a = off[it >= P ? it - P : it] + (it >= P ? n : 0);
        const uint32_t srel = (uint32_t)(v - a) << 8;
        ee.run(v, [&](int32_t k, int32_t dr, int32_t dc, int32_t sc) {
            edges[e++] = make_uint2(((uint32_t)(k - (int32_t)v) & 0xffffu) | ((uint32_t)dr << 16) | ((uint32_t)dc << 24),
                                    ((uint32_t)sc & 0xffu) | srel);
        });
    }));
#endif
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: chaining DP (collinearity.py:81-93 / :103-115), one work item = (pair, direction).
//
// The sweep is order-dependent (strict `<` first-writer-wins, and a `break` whose state changes only on
// updates), so the sources of an item are processed strictly in the reference's order by ONE executor
// thread; the state machine is the reference's, verbatim:
//   gap = mg2, row_old = row; for each candidate: if dc > gap and row_k beyond row_old: break;
//   if s > 0 and score_k < score_i + s: update, gap = min(dc, gap), row_old = row_k.
// What makes it fast is that nothing on the executor's dependency chain touches HBM or L2:
//  * scores and path lengths of the mg-bounded window live in a shared-memory ring indexed by the point's
//    position (a source only reaches points at most `span` positions ahead: the next mg1-1 rows are a
//    contiguous slice of the sorted arrays);
//  * a helper warp keeps the ring filled ahead of the executor (TMA bulk copies for the aligned interior,
//    plain coalesced loads at the ends), writes final scores back behind it, and pulls the executor's
//    read-only streams (edge offsets, edge records) into L1 ahead of use;
//  * pred / depth / used are store-only for the executor (fire and forget).
// Items are independent, so all (pair, direction) items of a comparison run concurrently, one CTA each,
// longest item first.  k_chain_generic is the same loop on global memory (oversized windows, emulation).
// ------------------------------------------------------------------------------------------------
struct ChainArgs {
    int64_t n;
    int32_t P, mg2, RS, span;
    const int32_t *order, *G;
    const int64_t *off;
    const uint32_t *eoff;
    const uint2 *edges;
    int32_t *S, *root;
    int2 *pd;
    uint8_t *used;
    long long *dbg;                  // optional per-item {rounds, cycles} (diagnostics), may be NULL
    // streaming export (icl_dp_begin): a finished item writes its reversed scores straight into the caller's
    // page-locked buffers and posts its id, so the host can sort it while heavier items are still running
    int32_t *hkey_fwd, *hkey_rev;    // device-visible host memory, NULL = no streaming
    volatile int32_t *hqueue;        // [items] completion flags in host memory
    const int32_t *imap;             // mirrored position - n -> forward position
};

// the reference sweep over the flat edge stream of one item, on global memory (oversized windows, emulation)
ICL_HD void chain_generic_item(const ChainArgs &A, int64_t j) {
    int32_t it = A.order[j];
    int32_t p = it >= A.P ? it - A.P : it;
    int64_t a = A.off[p] + (it >= A.P ? A.n : 0), b = A.off[p + 1] + (it >= A.P ? A.n : 0);
    if (a == b) return;
    int64_t v = a;
    int32_t Sv = A.S[v], dv = 1, rootv = (int32_t)v, gap = A.mg2, dr_old = 0;
    bool uv = false, broken = false;
    for (uint32_t e = A.eoff[a], e1 = A.eoff[b]; e < e1; e++) {
        uint2 rec = A.edges[e];
        const int64_t nv = a + (rec.y >> 8);
        if (nv != v) {                                           // next source that has candidates
            if (uv) A.used[v] = 1;
            v = nv;
            Sv = A.S[v];
            dv = (Sv != A.G[v]) ? A.pd[v].y : 1;
            rootv = (Sv != A.G[v]) ? A.root[v] : (int32_t)v;
            gap = A.mg2; dr_old = 0; uv = false; broken = false;
        }
        int32_t dc = rec.x >> 24, dr = (rec.x >> 16) & 255, sc = rec.y & 255;
        broken = broken || (dc > gap && dr > dr_old);            // collinearity.py:83-84 (break: stays broken)
        int64_t k = v + (rec.x & 0xffffu);
        if (!broken && sc > 0 && A.S[k] < Sv + sc) {              // :87 strict <
            A.S[k] = Sv + sc;
            A.pd[k] = make_int2((int32_t)v, dv + 1);
            A.root[k] = rootv;
            A.used[k] = 1;
            uv = true;
            gap = dc < gap ? dc : gap;                            // :92
            dr_old = dr;                                          // :93
        }
    }
    if (uv) A.used[v] = 1;
}

#ifndef ICL_EMU
#define ICL_EDGE_NCH 4               // edge ring: 4 chunks of 512 records = 16 KB
#define ICL_USED_FLAG 0x80000000u

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// One CTA per (pair, direction) item: warp 0 lane 0 is the executor, warp 1 the helper.
//   ring[RS]  : 16-byte slots {score, depth | used-as-source flag, pred, tree root} of the points inside the window
//   ering     : the item's edge records, streamed by TMA in 4 KB chunks (4-deep), consumed strictly in order
template <bool BATCH, bool DBG>
__global__ void __launch_bounds__(64) k_chain_ring(ChainArgs A) {
    extern __shared__ __align__(128) unsigned char dyn[];
    uint2 *ering = (uint2 *)dyn;                                           // ICL_EDGE_NCH * ICL_EDGE_CHUNK records
    int4 *ring = (int4 *)(dyn + ICL_EDGE_NCH * ICL_EDGE_CHUNK * sizeof(uint2));
    __shared__ volatile uint32_t ctl_loaded, ctl_retired, ctl_echunk;
    __shared__ __align__(8) uint64_t ebar[ICL_EDGE_NCH];
    const int32_t it = A.order[blockIdx.x];
    const int32_t p = it >= A.P ? it - A.P : it;
    const uint32_t a = (uint32_t)(A.off[p] + (it >= A.P ? A.n : 0)), b = (uint32_t)(A.off[p + 1] + (it >= A.P ? A.n : 0));
    if (a == b) return;
    const uint32_t e0 = A.eoff[a], e1 = A.eoff[b];
    const uint32_t c0 = e0 / ICL_EDGE_CHUNK;
    const uint32_t M = (uint32_t)A.RS - 1, EM = ICL_EDGE_NCH * ICL_EDGE_CHUNK - 1;
    if (threadIdx.x == 0) {
        ctl_loaded = a;
        ctl_retired = a;
        ctl_echunk = c0;
        for (int i = 0; i < ICL_EDGE_NCH; i++) mbar_init(&ebar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        if (BATCH) {
            // ------------------------------------------------------------ executor warp: 32 records per round
            // Lane l replays record e + l speculatively against the ring as it stands at the start of the round.
            // * a record names its source (srel), so every lane finds its operands with two shared-memory loads;
            // * the sweep state inside a source (gap, row_old, break — collinearity.py:81-93) is resolved with
            //   one segmented warp scan over the would-update flags (legal: SURVEY.md A.3, the break position only
            //   depends on the records before it);
            // * a lane is stale if an earlier lane of the round really updates the slot it read as its source or
            //   as its target.  Claim table: atomicMax of {round, 31 - lane} per ring slot — the round stamp makes
            //   older claims lose, so the table is never reset.  The round commits the records before the first
            //   stale lane — lane 0 is never stale — and restarts at that lane, so every record is evaluated
            //   exactly as in the reference's sequential order.
            uint32_t *claim = (uint32_t *)(ring + A.RS);
            const uint32_t FULLM = 0xffffffffu;
            const int32_t mg2 = A.mg2;
            for (uint32_t i = lane; i < (uint32_t)A.RS; i += 32) claim[i] = 0u;
            uint32_t loaded = a;
            while (loaded < b && loaded <= a + (uint32_t)A.span) loaded = ctl_loaded;
            __threadfence_block();
            __syncwarp();
            // sweep state of the source that is open when a round starts (all lanes hold the same values)
            uint32_t cur_v = a;
            int32_t cgap = mg2, cdr = 0;
            bool cbroken = false;
            uint32_t e = e0, ready = c0 * ICL_EDGE_CHUNK, pubchunk = c0, rounds = 0;
            const uint32_t upto = (2u << lane) - 1u;              // lanes 0..lane
            const uint32_t ringlim = (uint32_t)A.RS - 64u;
            const long long t_begin = DBG ? clock64() : 0;
            long long t_wait = 0;
            while (e < e1) {
                __syncwarp();                                     // commits of the previous round are visible
                ++rounds;
                const uint32_t bend = e + 32 < e1 ? e + 32 : e1;
                if (e / ICL_EDGE_CHUNK != pubchunk) {             // chunks before the one holding e are consumed
                    pubchunk = e / ICL_EDGE_CHUNK;
                    __threadfence_block();
                    if (lane == 0) ctl_echunk = pubchunk;
                }
                while (ready < bend) {                            // TMA bytes of the chunks this round touches
                    const uint32_t ch = ready / ICL_EDGE_CHUNK;
                    const long long tw = DBG ? clock64() : 0;
                    mbar_wait(&ebar[ch % ICL_EDGE_NCH], ((ch - c0) / ICL_EDGE_NCH) & 1u);
                    if (DBG) t_wait += clock64() - tw;
                    ready = (ch + 1) * ICL_EDGE_CHUNK;
                }
                bool valid = e + lane < bend;
                const uint2 rec = ering[(e + lane) & EM];
                const uint32_t ypre = ering[(e + lane - 1) & EM].y;            // previous record (lane 0: unused)
                const uint32_t v0 = a + (ering[e & EM].y >> 8);               // first source of the round (broadcast)
                const uint32_t v = a + (rec.y >> 8);
                const uint32_t k = v + (rec.x & 0xffffu);
                // operands: issued first, everything up to `up` overlaps their latency (re-read in the rare case
                // that the staging check below fails)
                const uint32_t vslot = v & M, slot = k & M;
                int4 sl = ring[vslot];
                int32_t tx = ring[slot].x;
                const int32_t dc = rec.x >> 24, dr = (rec.x >> 16) & 255, sc = rec.y & 255;
                // a round only reaches as far as the ring can hold behind its first source
                valid = valid && k < v0 + ringlim;
                const uint32_t kmax = __reduce_max_sync(FULLM, valid ? k : 0u);
                // segments = records of one source; head = lane that opens mine, -1: open since before the round
                const bool first = lane == 0 ? (v != cur_v) : ((rec.y ^ ypre) >> 8) != 0;
                const uint32_t firstmask = __ballot_sync(FULLM, first);
                const uint32_t below = firstmask & upto;
                const int head = below ? 31 - __clz(below) : -1;
                const int seg_start = head < 0 ? 0 : head;
                const int segpos = lane - seg_start;
                const bool need8 = __any_sync(FULLM, valid && segpos >= 8);
                if (loaded < b && loaded <= kmax) {               // the targets are not staged yet
                    __threadfence_block();
                    if (lane == 0) ctl_retired = v0;              // everything before the round's first source is final
                    const long long tw = DBG ? clock64() : 0;
                    do { loaded = ctl_loaded; } while (loaded < b && loaded <= kmax);
                    if (DBG) t_wait += clock64() - tw;
                    __threadfence_block();
                    sl = ring[vslot];
                    tx = ring[slot].x;
                }
                const int32_t Sv = sl.x, dv = sl.y & 0x7fffffff, rootv = sl.w;
                const int32_t cand = Sv + sc;
                const bool up = valid && sc > 0 && tx < cand;                // :87, ignoring the break for now
                // sweep state before each record: ONE segmented scan over the would-update flags, min(dc) and
                // max(dr) packed as two 16-bit halves {255 - dc, dr} combined with a per-halfword max; the steps
                // beyond the longest segment of the round are skipped (warp-uniform).  (Skipping per step — 1, 2, 4 — by
                // the longest segment of the round was measured and is slower: 0.41 -> 0.47 ms on C2; the nested
                // uniform branches cost more than the shuffles they save.)
                uint32_t pk = up ? (((uint32_t)(255 - dc) << 16) | (uint32_t)dr) : 0u;
                {
                    uint32_t po = __shfl_up_sync(FULLM, pk, 1);
                    if (segpos >= 1) pk = __vmaxu2(pk, po);
                    po = __shfl_up_sync(FULLM, pk, 2);
                    if (segpos >= 2) pk = __vmaxu2(pk, po);
                    po = __shfl_up_sync(FULLM, pk, 4);
                    if (segpos >= 4) pk = __vmaxu2(pk, po);
                    if (need8) {
                        po = __shfl_up_sync(FULLM, pk, 8);
                        if (segpos >= 8) pk = __vmaxu2(pk, po);
                        po = __shfl_up_sync(FULLM, pk, 16);
                        if (segpos >= 16) pk = __vmaxu2(pk, po);
                    }
                }
                uint32_t pe = __shfl_up_sync(FULLM, pk, 1);
                if (segpos < 1) pe = 0u;
                const int32_t gx = 255 - (int32_t)(pe >> 16), gy = (int32_t)(pe & 0xffffu);
                const int32_t ig = head < 0 ? cgap : mg2, idr = head < 0 ? cdr : 0;
                const int32_t gap_b = gx < ig ? gx : ig, dr_b = gy > idr ? gy : idr;
                const bool brk = valid && dc > gap_b && dr > dr_b;            // :83-84
                const uint32_t brkmask = __ballot_sync(FULLM, brk);
                const uint32_t segmask = upto & ~((1u << seg_start) - 1u);    // lanes seg_start..lane
                const bool dead = (brkmask & segmask) != 0 || (head < 0 && cbroken);
                const bool u = up && !dead;
                // stale lanes: an earlier lane of the round updates a slot this lane read
                const uint32_t stamp = (rounds << 5) | (uint32_t)(31 - lane);
                if (u) atomicMax(&claim[slot], stamp);
                __syncwarp();
                const uint32_t c2 = claim[slot], c1 = claim[vslot];
                const bool stale = (c2 > stamp) | (c1 > stamp);
                const uint32_t stop = __ballot_sync(FULLM, stale || !valid);
                const int f = stop ? __ffs(stop) - 1 : 32;        // records committed this round (>= 1)
                if (lane < f && u) {
                    ring[slot] = make_int4(cand, dv + 1, (int32_t)v, rootv);
                    ring[vslot].y = dv | (int32_t)ICL_USED_FLAG;  // used-as-source (same value from every lane of a source)
                }
                // the source that is open after the last committed record
                const int32_t gap_a = u ? (dc < gap_b ? dc : gap_b) : gap_b, dr_a = u ? dr : dr_b;
                const uint32_t st = __shfl_sync(FULLM, (uint32_t)gap_a | ((uint32_t)dr_a << 8) | ((uint32_t)dead << 16), f - 1);
                cgap = (int32_t)(st & 255u);
                cdr = (int32_t)((st >> 8) & 255u);
                cbroken = (st >> 16) != 0;
                cur_v = __shfl_sync(FULLM, v, f - 1);
                e += (uint32_t)f;
                if ((rounds & 7u) == 0) {                         // let the helper retire finished points
                    __threadfence_block();
                    if (lane == 0) ctl_retired = cur_v;
                }
            }
            __syncwarp();
            __threadfence_block();
            if (DBG && lane == 0) {
                A.dbg[4 * blockIdx.x] = rounds; A.dbg[4 * blockIdx.x + 1] = clock64() - t_begin;
                A.dbg[4 * blockIdx.x + 2] = t_wait; A.dbg[4 * blockIdx.x + 3] = e1 - e0;
            }
            if (lane == 0) ctl_retired = b;
            return;
        }
        // ---------------------------------------------------------------- executor (one thread)
        if (lane != 0) return;
        uint32_t loaded = a, v = a;
        const int32_t mg2 = A.mg2;
        const uint32_t span = (uint32_t)A.span;
        // the window of the current source ([v, v + span]) is staged before any of its records is replayed,
        // so the record loop itself never looks at the helper's progress
        while (loaded < b && loaded <= v + span) loaded = ctl_loaded;
        __threadfence_block();
        int4 sl = ring[v & M];
        int32_t Sv = sl.x, dv1 = sl.y + 1, rootv = sl.w, gap = mg2, dr_old = 0;
        uint32_t uv = 0;
        bool broken = false;
        int32_t pub = 32;
        uint32_t e = e0;
        while (e < e1) {
            // ---- enter the chunk that holds record e: free the previous ones, wait for the TMA bytes
            const uint32_t ch = e / ICL_EDGE_CHUNK;
            __threadfence_block();
            ctl_echunk = ch;
            mbar_wait(&ebar[ch % ICL_EDGE_NCH], ((ch - c0) / ICL_EDGE_NCH) & 1u);
            uint32_t cend = (ch + 1) * ICL_EDGE_CHUNK;
            cend = cend < e1 ? cend : e1;
            uint2 nxt = ering[e & EM];
            for (; e < cend; e++) {
                const uint2 rec = nxt;
                nxt = ering[(e + 1) & EM];                        // next record early (unused at a chunk end)
                const uint32_t nv = a + (rec.y >> 8);
                if (nv != v) {                                    // next source that has candidates
                    ring[v & M].y = (dv1 - 1) | (int32_t)uv;      // used-as-source flag of the finished source
                    v = nv;
                    if (loaded < b && loaded <= v + span) {       // the new source's window is not staged yet:
                        __threadfence_block();                    // retire everything before v so the helper can move on
                        ctl_retired = v;
                        pub = 32;
                        do { loaded = ctl_loaded; } while (loaded < b && loaded <= v + span);
                        __threadfence_block();
                    }
                    sl = ring[v & M];
                    Sv = sl.x; dv1 = sl.y + 1; rootv = sl.w;
                    gap = mg2; dr_old = 0; uv = 0; broken = false;
                    if (--pub == 0) {                             // let the helper retire finished points
                        pub = 32;
                        __threadfence_block();
                        ctl_retired = v;
                    }
                }
                const int32_t dc = rec.x >> 24, dr = (rec.x >> 16) & 255, sc = rec.y & 255;
                broken = broken || (dc > gap && dr > dr_old);     // collinearity.py:83-84
                const uint32_t slot = (v + (rec.x & 0xffffu)) & M;
                const int32_t cand = Sv + sc;
                if (!broken && sc > 0 && ring[slot].x < cand) {   // :87 strict <
                    ring[slot] = make_int4(cand, dv1, (int32_t)v, rootv);
                    uv = ICL_USED_FLAG;
                    gap = dc < gap ? dc : gap;                    // :92
                    dr_old = dr;                                  // :93
                }
            }
        }
        ring[v & M].y = (dv1 - 1) | (int32_t)uv;
        __threadfence_block();
        ctl_retired = b;
        return;
    }
    // -------------------------------------------------------------------- helper warp
    // streamed export (icl_dp_begin): the score of every retired point also goes straight to the caller's page-locked
    // buffer, at its place in the REVERSED (loc1, loc2) order of the pair — the array pandas' nargsort(ascending=False)
    // hands to numpy.argsort.  Forward items: 32 consecutive points are 128 contiguous bytes; mirrored items visit the
    // rows in the same order as the reversed forward order and only run the other way inside a row, so a warp's 32
    // stores still cover one contiguous range (up to the rows cut at its ends).
    const bool mirrored = it >= A.P;
    int32_t *hk = (BATCH && A.hqueue) ? (mirrored ? A.hkey_rev : A.hkey_fwd) : nullptr;
    const uint32_t hbase = (uint32_t)A.off[p] + (uint32_t)A.off[p + 1] - 1u;
    uint32_t wb = a, ld = a;
    uint32_t next_chunk = c0;
    const uint32_t last_chunk = e1 > e0 ? (e1 - 1) / ICL_EDGE_CHUNK : c0;
    const bool has_edges = e1 > e0;
    while (wb < b) {
        uint32_t r = 0, ech = 0;
        if (lane == 0) { r = ctl_retired; ech = ctl_echunk; }
        r = __shfl_sync(0xffffffffu, r, 0);
        ech = __shfl_sync(0xffffffffu, ech, 0);
        __threadfence_block();
        bool progress = false;
        // 1. edge chunks: keep the 4-deep ring full (chunks < ech are consumed)
        if (has_edges && lane == 0) {
            while (next_chunk <= last_chunk && next_chunk < ech + ICL_EDGE_NCH) {
                uint64_t *bar = &ebar[next_chunk % ICL_EDGE_NCH];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // executor's reads of the slot -> TMA write
                mbar_expect_tx(bar, ICL_EDGE_CHUNK * (uint32_t)sizeof(uint2));
                tma_load_1d(ering + (next_chunk % ICL_EDGE_NCH) * ICL_EDGE_CHUNK, A.edges + (size_t)next_chunk * ICL_EDGE_CHUNK,
                            ICL_EDGE_CHUNK * (uint32_t)sizeof(uint2), bar);
                next_chunk++;
            }
        }
        next_chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
        // 2. results of retired points -> HBM (coalesced): score, pred, depth, used
        if (r > ld) r = ld;
        if (r > wb) {
            for (uint32_t v = wb + lane; v < r; v += 32) {
                const int4 sl = ring[v & M];
                const int32_t g = A.G[v];
                A.S[v] = sl.x;
                A.pd[v] = (sl.x != g) ? make_int2(sl.z, sl.y & 0x7fffffff) : make_int2(-1, 1);
                A.root[v] = sl.w;
                A.used[v] = (sl.x != g) || (sl.y < 0);
                if (hk) hk[hbase - (mirrored ? (uint32_t)__ldg(A.imap + (v - (uint32_t)A.n)) : v)] = sl.x;
            }
            wb = r;
            progress = true;
        }
        // 3. stage the next points of the window: score = grading, path length 1 (get_matrix :35-37)
        uint32_t lim = wb + (uint32_t)A.RS;
        lim = lim < b ? lim : b;
        if (ld < lim) {
            __syncwarp();
            for (uint32_t v = ld + lane; v < lim; v += 32) ring[v & M] = make_int4(A.G[v], 1, -1, (int32_t)v);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) ctl_loaded = lim;
            ld = lim;
            progress = true;
        }
        if (!progress) __nanosleep(100);
    }
    if (hk) {                                         // every score of the item is on its way: raise the completion flag
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            A.hqueue[it] = 1;
            __threadfence_system();
        }
    }
}
#endif


static int phase_chain(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->S.reserve(st, V) || c->pd.reserve(st, V) || c->used.reserve(st, V) || c->root.reserve(st, V)) return st.code;
    ChainArgs A;
    A.n = n; A.P = P; A.mg2 = c->prm.mg2; A.RS = 0; A.span = c->max_span;
    A.order = c->d_item_order.p; A.G = c->G.p; A.off = c->d_pair_off.p; A.eoff = c->eoff.p; A.edges = c->edges.p;
    A.S = c->S.p; A.pd = c->pd.p; A.used = c->used.p; A.root = c->root.p;
    A.dbg = nullptr;
    A.hkey_fwd = A.hkey_rev = nullptr; A.hqueue = nullptr; A.imap = c->imap.p;
    c->streamed = false;
    if (getenv("ICL_B200_CHAIN_DEBUG")) {
        if (c->chain_dbg.reserve(st, 8 * (size_t)P + 8)) return st.code;
        ICL_CUDA(st, dev_memset(c->chain_dbg.p, 0, (8 * (size_t)P + 8) * sizeof(long long), s));
        A.dbg = c->chain_dbg.p;
    }
    if (P == 0 || n == 0) return 0;
#ifndef ICL_EMU
    if (c->chain_mode != 0) {
        // ---- ring kernels: window span + what a round / the executor may run ahead of the last retire point
        int64_t want = (int64_t)c->max_span + 256;
        int64_t RS = 512;
        while (RS < want) RS <<= 1;
        if (RS <= 8192) {
            A.RS = (int32_t)RS;
            const bool batch = c->chain_mode == 1;
            if (batch && c->sq_queue) {                            // icl_dp_begin: items export themselves as they finish
                A.hkey_fwd = c->sq_fwd; A.hkey_rev = c->sq_rev; A.hqueue = c->sq_queue;
                c->streamed = true;
            }
            size_t smem = (size_t)RS * sizeof(int4) + (size_t)ICL_EDGE_NCH * ICL_EDGE_CHUNK * sizeof(uint2) +
                          (batch ? (size_t)RS * sizeof(uint32_t) : 0);
            auto launch = [&](void (*kern)(ChainArgs)) {
                ICL_CUDA(st, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                kern<<<(unsigned)(2 * P), 64, smem, s>>>(A);
                ICL_COUNT_KERNEL();
                return 0;
            };
            if (launch(batch ? (A.dbg ? k_chain_ring<true, true> : k_chain_ring<true, false>) : k_chain_ring<false, false>)) return st.code;
            ICL_CUDA(st, cudaGetLastError());
            c->chain_kernel = batch ? 1 : 3;
            return 0;
        }
    }
#endif
    ICL_CUDA(st, d2d(c->S.p, c->G.p, (size_t)V * sizeof(int32_t), s));            // get_matrix: score = grading
    ICL_CUDA(st, dev_memset(c->used.p, 0, (size_t)V, s));
    ICL_CUDA(st, foreach_warp(s, 2 * (int64_t)P, ICL_LAMBDA(int64_t j) { chain_generic_item(A, j); }));
    {
        const int32_t *S = c->S.p, *G = c->G.p;
        int2 *pd = c->pd.p;
        int32_t *root = c->root.p;
        ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) { if (S[v] == G[v]) { pd[v] = make_int2(-1, 1); root[v] = (int32_t)v; } }));
    }
    c->chain_kernel = 0;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: icl_dp
// ------------------------------------------------------------------------------------------------
// diagnostics: per-item {rounds, cycles, wait cycles, records} of the DP kernel (env ICL_B200_CHAIN_DEBUG = file)
static int dump_chain_debug(icl_ctx *c) {
    Status &st = c->st;
    cudaStream_t s = c->stream;
    if (getenv("ICL_B200_CHAIN_DEBUG") && c->chain_dbg.p && c->P > 0) {
        std::vector<long long> h(8 * (size_t)c->P);
        ICL_CUDA(st, d2h(h.data(), c->chain_dbg.p, h.size() * sizeof(long long), s));
        ICL_CUDA(st, dev_sync(s));
        FILE *f = fopen(getenv("ICL_B200_CHAIN_DEBUG"), "w");
        if (f) {
            for (int32_t j = 0; j < 2 * c->P; j++) {
                int32_t it = c->h_item_order[j], p = it >= c->P ? it - c->P : it;
                fprintf(f, "%d %d %lld %d %lld %lld %lld %lld\n", j, it, (long long)(c->h_pair_off[p + 1] - c->h_pair_off[p]), c->h_nrows[p], h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
            }
            fclose(f);
        }
    }
    return 0;
}

// Everything icl_dp does.  With `queue` (icl_dp_begin) and the shared-memory ring kernel, finished items write
// their scores to the caller's buffers themselves and the call returns right after the launch; otherwise the
// scores are exported and copied after the DP and all completion flags (if any) are raised at the end.
static int dp_run(icl_ctx *c, const icl_params *prm, int32_t *score_fwd, int32_t *score_rev, int32_t *queue) {
    if (!c) return ICL_ERR_ARG;
    if (!c->uploaded) return fail(c, ICL_ERR_STATE, "icl_dp: icl_upload has not succeeded");
    if (!prm || (c->n > 0 && (!score_fwd != !score_rev)) || (queue && c->n > 0 && !score_fwd))
        return fail(c, ICL_ERR_ARG, "icl_dp: null argument");
    const bool want_keys = score_fwd != nullptr;               // NULL, NULL: the caller ranks on the device (icl_rank_device)
    if (prm->mg1 < 2 || prm->mg1 > 255 || prm->mg2 < 2 || prm->mg2 > 255)
        return fail(c, ICL_ERR_LIMIT, "icl_dp: mg values must be in 2..255");
    if (prm->g0 <= 0 || prm->over_gap < 1) return fail(c, ICL_ERR_ARG, "icl_dp: grading[0] and over_gap must be positive");
    if (c->gtab.n > 0 && c->n > 0 && c->max_grad >= c->gtab.n)
        return fail(c, ICL_ERR_ARG, "icl_dp: a grading table is set but the upload holds a class id outside it");
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    LaunchScope scope(c, 0);
    c->dp_done = c->dp_begun = false;
    c->order_on_device = false;
    c->prm = *prm;
    c->sq_fwd = score_fwd; c->sq_rev = score_rev; c->sq_queue = queue;
    if (queue) for (int64_t i = 0; i < 2 * (int64_t)c->P; i++) queue[i] = 0;
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
    c->counters[0] = c->n_edges;
    c->counters[3] = c->R;
    c->counters_ext[0] = c->max_span;
    c->counters_ext[1] = c->chain_kernel;
    if (c->streamed) {                                         // the kernel exports; icl_dp_end waits for it
        for (int32_t p = 0; p < c->P; p++)                     // empty items never run: nothing to wait for
            if (c->h_pair_off[p + 1] == c->h_pair_off[p]) queue[p] = queue[c->P + p] = 1;
        c->timers[ICL_T_EXPORT].start(s); c->timers[ICL_T_EXPORT].stop(s);
        c->timers[ICL_T_D2H].start(s); c->timers[ICL_T_D2H].stop(s);
        c->dp_begun = true;
        return ICL_OK;
    }
    c->timers[ICL_T_EXPORT].start(s);
    if (want_keys && c->score32.reserve(st, 2 * n)) return st.code;
    if (want_keys) {
        // per pair segment in REVERSED point order: the array pandas' nargsort(ascending=False) hands to numpy
        const int32_t *S = c->S.p, *mmap = c->mmap.p, *ps = c->pair_s.p;
        const int64_t *off = c->d_pair_off.p;
        int32_t *o = c->score32.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            int32_t p = ps[i];
            int64_t src = off[p] + (off[p + 1] - 1 - i);
            o[i] = S[src];
            o[n + i] = S[mmap[src]];
        }));
    }
    c->timers[ICL_T_EXPORT].stop(s);
    c->timers[ICL_T_D2H].start(s);
    if (n > 0 && want_keys) {
        if (score_rev == score_fwd + n) {                      // one buffer of 2n: a single transfer
            ICL_CUDA(st, d2h(score_fwd, c->score32.p, (size_t)2 * n * sizeof(int32_t), s));
        } else {
            ICL_CUDA(st, d2h(score_fwd, c->score32.p, (size_t)n * sizeof(int32_t), s));
            ICL_CUDA(st, d2h(score_rev, c->score32.p + n, (size_t)n * sizeof(int32_t), s));
        }
    }
    c->timers[ICL_T_D2H].stop(s);
    // with scores for the host the call ends when they have arrived; without (device rank), the next phase is simply
    // queued behind this one on the stream
    if (want_keys || queue || getenv("ICL_B200_CHAIN_DEBUG")) ICL_CUDA(st, dev_sync(s));
    if (dump_chain_debug(c)) return st.code;
    if (queue) for (int64_t i = 0; i < 2 * (int64_t)c->P; i++) queue[i] = 1;     // not streamed: everything is complete now
    c->dp_begun = true;
    c->dp_done = queue == nullptr;                             // icl_dp_begin: only icl_dp_end completes the phase
    return ICL_OK;
}

extern "C" int icl_dp(icl_ctx *c, const icl_params *prm, int32_t *score_fwd, int32_t *score_rev) {
    return dp_run(c, prm, score_fwd, score_rev, nullptr);
}

extern "C" int icl_dp_begin(icl_ctx *c, const icl_params *prm, int32_t *key_fwd, int32_t *key_rev, int32_t *done_flags) {
    if (c && c->P > 0 && !done_flags) return fail(c, ICL_ERR_ARG, "icl_dp_begin: null done_flags");
    return dp_run(c, prm, key_fwd, key_rev, done_flags);
}

extern "C" int icl_dp_end(icl_ctx *c) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_begun) return fail(c, ICL_ERR_STATE, "icl_dp_end: icl_dp_begin has not succeeded");
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    ICL_CUDA(st, dev_sync(c->stream));
    if (c->streamed && dump_chain_debug(c)) return st.code;
    c->dp_done = true;
    return ICL_OK;
}

extern "C" int icl_layout(icl_ctx *c, int64_t *pair_off, int32_t *perm) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_begun) return fail(c, ICL_ERR_STATE, "icl_layout: icl_dp has not succeeded");
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
        const int32_t *S = c->S.p, *mmap = c->mmap.p, *imap = c->imap.p;
        const int2 *pd = c->pd.p;
        const uint8_t *us = c->used.p;
        int32_t *o1 = t1.p, *o2 = t2.p, *o3 = t3.p;
        uint8_t *o4 = t4.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            int64_t v = dir ? mmap[i] : i;
            o1[i] = S[v];
            int32_t q = pd[v].x;
            o2[i] = (q < 0) ? -1 : (dir ? imap[q - n] : q);
            o3[i] = pd[v].y;
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
// phase: extraction (collinearity.py:49-68 with max_path :117-144).
//
// Reference semantics: tops are visited in descending-score order (the caller's numpy permutation); a top
// that is still present is either accepted together with the present part of its path
// (L >= over_gap and m / L > coverage_ratio, :129) or removed alone (:143).
//
// Two facts make this parallel without changing a single decision:
//  * The removed members of a path always form a root-side prefix: an ancestor has a strictly smaller score
//    than its descendants (every DP step adds s > 0), so it is examined after them and can only have been
//    removed as part of an accepted block — which removes every present ancestor as well.  Hence the present
//    members of a path are the walk from the top up to the first removed node.
//  * A decision only reads and writes `present` flags of the top's own predecessor tree.  Trees are therefore
//    independent; only the ORDER in which accepted blocks are emitted (and bump `times`, phase B) is global.
// So: key every used point by (tree root, rank position), radix sort, let one thread run each tree in rank
// order, record accepted blocks at their rank position, and recover the reference's acceptance order with a
// prefix sum over rank positions.
// ------------------------------------------------------------------------------------------------
#define ICL_TREE_SMALL 32           // trees of at most this many used points: one thread each, on global memory
#define ICL_TREE_CAP 2048            // larger ones up to this size: one warp each, staged in shared memory
#define ICL_TREE_HUGE 256            // ... the ones above this size in a launch of their own (a power of two, like the cap:
                                     // the hash of a tree has up to 2 * cap slots)
#define ICL_TREE_SMEM(cap) ((cap) * 45 + 64)
// One predecessor tree on global memory: its points are the sorted entries [t0, uend), in rank order.
struct TreeWalk {
    int64_t n, V;
    const uint32_t *key, *tpos;
    const int32_t *rank2v;
    const int2 *pd;
    const int32_t *imap, *l1, *l2;
    uint8_t *present;
    int2 *hint;
    uint32_t *acc_m, *acc_ms;
    int32_t *acc_L, *mscr;
    int4 *acc_box;
    int32_t over_gap;
    double cov;
    // end of the run of equal keys that starts at t0: gallop, then bisect
    ICL_HD int64_t run_end(int64_t t0) const {
        const uint32_t rt = key[t0];
        int64_t lo = t0 + 1, step = 1;
        while (lo < V && key[lo] == rt) { lo += step; step <<= 1; }
        int64_t a0 = lo - (step >> 1), b0 = lo < V ? lo : V;             // key[a0 - 1] == rt (or a0 == t0 + 1), key[b0] != rt
        if (a0 < t0 + 1) a0 = t0 + 1;
        while (a0 < b0) {
            const int64_t mid = a0 + ((b0 - a0) >> 1);
            if (key[mid] == rt) a0 = mid + 1; else b0 = mid;
        }
        return a0;
    }
    ICL_HD void walk(int64_t t0, int64_t uend) const {
        int32_t ver = 0;                                                    // accepted blocks of this tree so far
        uint32_t cursor = (uint32_t)t0;                                     // member scratch: the tree's own slice
        // The tops of a tree are examined strictly in rank order, but everything a top needs except `present` and
        // the memo is read-only here, and those two only change for OTHER tops when a block is accepted (ver++).
        // So the loads run three tops ahead of the decisions (rank position at distance 3, point at 2, node record
        // + present + memo at 1) and a prefetched present / memo is re-read only if a block was accepted since.
        // A memo fetched before a rejected child wrote it is merely older: it is used only when its version
        // matches, otherwise the path is walked — both are exact.
        uint32_t tA = tpos[t0];
        uint32_t tB = t0 + 1 < uend ? tpos[t0 + 1] : 0u, tC = t0 + 2 < uend ? tpos[t0 + 2] : 0u, tD = t0 + 3 < uend ? tpos[t0 + 3] : 0u;
        int32_t topA = rank2v[tA], topB = t0 + 1 < uend ? rank2v[tB] : 0, topC = t0 + 2 < uend ? rank2v[tC] : 0;
        int2 ndA = ICL_LDG(pd + topA), hA = hint[topA];
        uint8_t prA = present[topA];
        int32_t verA = 0, verB = 0;
        int2 ndB = make_int2(0, 0), hB = make_int2(0, 0);
        uint8_t prB = 0;
        if (t0 + 1 < uend) { ndB = ICL_LDG(pd + topB); hB = hint[topB]; prB = present[topB]; }
        for (int64_t u = t0; u < uend; u++) {                               // the tree's points in rank order
            const uint32_t tE = u + 4 < uend ? tpos[u + 4] : 0u;
            const int32_t topD = u + 3 < uend ? rank2v[tD] : 0;
            int2 ndC = make_int2(0, 0), hC = make_int2(0, 0);
            uint8_t prC = 0;
            const int32_t verC = ver;
            if (u + 2 < uend) { ndC = ICL_LDG(pd + topC); hC = hint[topC]; prC = present[topC]; }
            const uint32_t t = tA;
            const int32_t top = topA;
            const int2 nd = ndA;
            uint8_t pr = prA;
            int2 h = hA;
            if (verA != ver) { pr = present[top]; h = hint[top]; }
            if (pr) {
                const int32_t L = nd.y;
                int32_t m, q;
                if (ver == 0) {
                    m = L;                                                  // nothing of this tree removed yet
                } else if (h.y == ver) {
                    m = L - h.x;                                            // boundary known from a rejected child
                } else {
                    m = 1; q = nd.x;
                    while (q >= 0 && present[q]) { m++; q = ICL_LDG(pd + q).x; }     // present part of the path
                }
                if (L >= over_gap && (double)m / (double)L > cov) {         // :129
                    // remove the members (:130), list them, bounding box (:131)
                    const bool rev = top >= n;
                    int32_t l2min = 0x7fffffff, l2max = -1, ifirst = 0, ilast = 0;
                    q = top;
                    for (int32_t k = 0; k < m; k++) {
                        if (q < 0) break;                                   // (only with a caller's broken order, reported below)
                        present[q] = 0;
                        const int32_t i = rev ? imap[q - n] : q;            // forward sorted position
                        mscr[cursor + k] = i;
                        const int32_t yv = l2[i];
                        l2min = yv < l2min ? yv : l2min;
                        l2max = yv > l2max ? yv : l2max;
                        if (k == 0) ifirst = i;
                        ilast = i;
                        q = ICL_LDG(pd + q).x;
                    }
                    ver++;                                                  // removed ancestors changed: memos expire
                    const int32_t la = l1[ifirst], lb = l1[ilast];
                    acc_m[t] = (uint32_t)m;
                    acc_L[t] = L;
                    acc_ms[t] = cursor;
                    acc_box[t] = make_int4(la < lb ? la : lb, la < lb ? lb : la, l2min, l2max);
                    cursor += (uint32_t)m;
                } else {
                    present[top] = 0;                                       // :143
                    // the parent sees the same removed ancestor until the next accepted block of this tree;
                    // without the memo a long path whose root side was taken by another block costs O(L^2) hops
                    if (m > 1) hint[nd.x] = make_int2(L - m, ver);
                }
            }
            tA = tB; topA = topB; ndA = ndB; hA = hB; prA = prB; verA = verB;
            tB = tC; topB = topC; ndB = ndC; hB = hC; prB = prC; verB = verC;
            tC = tD; topC = topD; tD = tE;
        }
    }
};

#ifndef ICL_EMU
// heads of the sorted runs (one predecessor tree each) -> the queue of short trees or of long ones.  The queue of short
// trees is pushed once per CTA and 256 entries (one atomic on the shared cursor per warp was the whole kernel: 119 k
// same-address atomics, 65 us on C2).
__global__ void __launch_bounds__(256) k_tree_heads(TreeWalk W, int2 *smallq, int2 *bigq, int64_t bigcap, int32_t *stats) {
    __shared__ int32_t wcnt[8], cbase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < W.V; base += stride) {       // (CTA-uniform trip count)
        const int64_t t0 = base + threadIdx.x;
        // a run ends where the key changes: most runs (a tree holds a few points) end inside the warp's 32 entries, where the
        // ballot of the changes shows the end; only a run that reaches past them is searched (gallop + bisect)
        const uint32_t rt = t0 < W.V ? W.key[t0] : 0xffffffffu;
        uint32_t prev = __shfl_up_sync(0xffffffffu, rt, 1);
        if (lane == 0) prev = t0 > 0 ? W.key[t0 - 1] : 0xffffffffu;
        const bool change = t0 <= W.V && (t0 == 0 || t0 == W.V || prev != rt);
        const uint32_t cm = __ballot_sync(0xffffffffu, change);
        bool head = t0 < W.V && change && rt < (uint32_t)W.V;
        int64_t uend = 0;
        if (head) {
            const uint32_t later = lane < 31 ? cm >> (lane + 1) : 0u;
            uend = later ? t0 + __ffs(later) : W.run_end(t0);
        }
        // a path is made of points of its tree: a tree with fewer used points than over_gap cannot yield a block (:129), and
        // nothing else looks at its `present` flags — most trees are such pairs of noise hits
        head = head && uend - t0 >= W.over_gap;
        const bool big = head && uend - t0 > ICL_TREE_SMALL && uend - t0 <= ICL_TREE_CAP;
        const bool small = head && !big;
        const uint32_t ms = __ballot_sync(0xffffffffu, small);
        if (lane == 0) wcnt[warp] = __popc(ms);
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t tot = 0;
            for (int w = 0; w < 8; w++) { const int32_t k = wcnt[w]; wcnt[w] = tot; tot += k; }
            cbase = tot ? atomicAdd(&stats[3], tot) : 0;
        }
        __syncthreads();
        if (small) smallq[cbase + wcnt[warp] + __popc(ms & ((1u << lane) - 1u))] = make_int2((int32_t)t0, (int32_t)uend);
        // the staged walk lasts as long as its largest tree: the long ones queue from the far end and are taken first
        if (big) {
            if (uend - t0 > ICL_TREE_HUGE) bigq[bigcap - 1 - atomicAdd(&stats[4], 1)] = make_int2((int32_t)t0, (int32_t)uend);
            else bigq[atomicAdd(&stats[1], 1)] = make_int2((int32_t)t0, (int32_t)uend);
        }
    }
}

// one thread per short tree, on global memory
__global__ void __launch_bounds__(128) k_tree_walk_small(TreeWalk W, const int2 *smallq, const int32_t *stats) {
    const int32_t count = stats[3];
    for (int32_t i = blockIdx.x * 128 + threadIdx.x; i < count; i += gridDim.x * 128) {
        const int2 rg = smallq[i];
        W.walk(rg.x, rg.y);
    }
}
#endif

#ifndef ICL_EMU
struct TreeArgs {
    int64_t n, V;
    const uint32_t *key, *tpos;      // sorted tree keys; rank position of every sorted entry
    const int32_t *rank2v;
    const int2 *pd;
    const int32_t *imap, *l1, *l2;
    const int2 *big;                 // queue of [t0, uend): the longest trees from the far end (slot bigcap - 1 downwards), the others from slot 0
    int64_t bigcap;
    int32_t *stats;                  // [0] / [5] queue heads (the two launches), [1] / [4] queue lengths (written by the tree-head pass)
    uint32_t *acc_m, *acc_ms;
    int32_t *acc_L, *mscr;
    int4 *acc_box;
    int32_t over_gap;
    double cov;
};

// The extraction of one predecessor tree (same decisions, same outputs as the thread-per-tree walk above) with the tree
// in shared memory: the warp gathers its points once — rank position, forward position, loc2, path length, and the
// predecessor as an index INTO the tree (through a hash of the tree's points) — and lane 0 then walks tops, paths and
// members at shared-memory latency instead of one dependent DRAM round trip per hop.
// Two launches share the code: trees of up to ICL_TREE_HUGE points (11.6 KB of shared memory per warp: the mass of the
// staged points, spread over 16 warps per SM) and the few longer ones (92 KB per warp, 2 per SM), side by side.  One
// launch with the large footprint for all of them ran 296 warps for ~1000 points each: 130 us on C2 whatever the
// largest tree cost.
template <int CAP, bool HUGE>
__global__ void __launch_bounds__(32) k_tree_walk_staged(TreeArgs T) {
    extern __shared__ __align__(16) unsigned char tsm[];
    int32_t *s_t = (int32_t *)tsm;                             // rank position
    int32_t *s_f = s_t + CAP;                         // forward sorted position of the point
    int32_t *s_y = s_f + CAP;                         // loc2
    int32_t *s_pl = s_y + CAP;                        // predecessor: index into the tree, -1 for the root
    int32_t *s_L = s_pl + CAP;                        // path length
    int32_t *s_x = s_L + CAP;                         // loc1
    int32_t *s_mm = s_x + CAP;                        // per path length L <= N: the smallest m with (double) m / L > cov
    int2 *s_h = (int2 *)(s_mm + CAP);                  // 2 * CAP slots: hash {point, index}, later the memo {depth, version}
    uint8_t *s_pres = (uint8_t *)(s_h + 2 * CAP);
    const int lane = threadIdx.x;
    const uint32_t FULL = 0xffffffffu;
    for (;;) {
        int32_t qi = 0;
        if (lane == 0) qi = atomicAdd(&T.stats[HUGE ? 5 : 0], 1);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= T.stats[HUGE ? 4 : 1]) break;
        const int2 rg = HUGE ? T.big[T.bigcap - 1 - qi] : T.big[qi];
        const int32_t t0 = rg.x, N = rg.y - rg.x;
        int32_t hsz = 64, hshift = 26;
        while (hsz < 2 * N) { hsz <<= 1; hshift--; }
        for (int32_t i = lane; i < hsz; i += 32) s_h[i] = make_int2(-1, -1);
        __syncwarp();
        // Gather of the tree, one level of the dependent chain (rank position -> point -> node record -> forward position
        // -> coordinates) per pass, eight loads of a lane in flight at a time (ICL_TREE_GATHER: registers first, stores after
        // — left as a plain loop the compiler keeps ONE load per lane in flight: 95 k cycles for the 678-point tree of C2;
        // the round of two dependent 5-deep chains per lane before it: 49 k).
        bool rev = false;
#define ICL_TREE_GATHER(TYPE, LOAD, STORE)                                                                   \
        for (int32_t j0 = lane; j0 < N; j0 += 256) {                                                        \
            TYPE r_[8];                                                                                      \
            _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) { const int32_t j = j0 + 32 * k_; if (j < N) { r_[k_] = LOAD; } } \
            _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) { const int32_t j = j0 + 32 * k_; if (j < N) { const TYPE r = r_[k_]; STORE; } } \
        }
        ICL_TREE_GATHER(int32_t, (int32_t)T.tpos[t0 + j], s_t[j] = r)
        ICL_TREE_GATHER(int32_t, T.rank2v[s_t[j]], s_x[j] = r)                              // (s_x: the point, until the last pass)
        ICL_TREE_GATHER(int2, T.pd[s_x[j]], (s_pl[j] = r.x, s_L[j] = r.y))
        ICL_TREE_GATHER(int32_t, (s_x[j] >= T.n ? T.imap[s_x[j] - T.n] : s_x[j]), s_f[j] = r)
        ICL_TREE_GATHER(int32_t, T.l2[s_f[j]], s_y[j] = r)
        if (lane == 0) rev = s_x[0] >= T.n;
        for (int32_t j = lane; j < N; j += 32) {                   // hash of the tree's points (before s_x is overwritten)
            const int32_t v = s_x[j];
            uint32_t h = ((uint32_t)v * 2654435761u) >> hshift;    // (the HIGH bits of the product: the low ones repeat with the key's)
            int32_t tries = 0;                                 // (bounded: a caller's broken order must not hang the GPU)
            while (atomicCAS(&s_h[h].x, -1, v) != -1 && ++tries < hsz) h = (h + 1) & (uint32_t)(hsz - 1);
            if (tries < hsz) s_h[h].y = j;
        }
        ICL_TREE_GATHER(int32_t, T.l1[s_f[j]], (s_x[j] = r, s_pres[j] = 1))
#undef ICL_TREE_GATHER
        for (int32_t j = N + lane; j < ((N + 15) & ~15); j += 32) s_pres[j] = 0;      // (the walk reads `present` 16 points at a time)
        rev = __shfl_sync(FULL, rev, 0);
        __syncwarp();
        for (int32_t j = lane; j < N; j += 32) {
            const int32_t p = s_pl[j];
            int32_t loc = -1;
            if (p >= 0) {                                      // the predecessor is a used point of the same tree
                uint32_t h = ((uint32_t)p * 2654435761u) >> hshift;
                int32_t tries = 0;
                while (s_h[h].x != p && s_h[h].x != -1 && ++tries < hsz) h = (h + 1) & (uint32_t)(hsz - 1);
                loc = s_h[h].x == p ? s_h[h].y : -1;
            }
            s_pl[j] = loc;
        }
        __syncwarp();
        for (int32_t j = lane; j < N; j += 32) s_h[j] = make_int2(-1, -1);   // the memo takes the hash's place
        // the coverage test `len(rem) / L > coverage_ratio` (:129) is a correctly rounded binary64 division in the reference;
        // the quotient grows with m, so per L there is a smallest m that passes — found with that very division, once per
        // length instead of once per examined top (a path of this tree is at most N long)
        for (int32_t L = lane + 1; L <= N; L += 32) {
            int32_t m = (int32_t)(T.cov * (double)L);
            m = m < 0 ? 0 : m > L ? L : m;
            bool stepped = false;                              // (m - 1 is then known to fail: one division less, of ~260 cycles each)
            while (m <= L && !((double)m / (double)L > T.cov)) { m++; stepped = true; }
            while (!stepped && m > 0 && (double)(m - 1) / (double)L > T.cov) m--;
            s_mm[L - 1] = m;                                   // L + 1 when no m <= L passes (coverage_ratio >= 1)
        }
        __syncwarp();
        if (lane == 0) {
            int32_t ver = 0;                                   // accepted blocks of this tree so far
            uint32_t cursor = (uint32_t)t0;                    // member scratch: the tree's own slice
            // The walk of a tree is sequential and the largest tree's time is the kernel's.  By the time their turn as a top
            // comes most points are gone — taken as members of an earlier block (the 678-point tree of C2: one block of 546)
            // — and a loop iteration per point cost 140 cycles each: the next PRESENT point is found 16 flags per load.
            for (int32_t u = 0; u < N; u++) {                  // the tree's points in rank order
                {
                    const int32_t base = u & ~15, k = u & 15;
                    const uint4 w = *(const uint4 *)(s_pres + base);
                    unsigned long long lo = (unsigned long long)w.x | ((unsigned long long)w.y << 32);
                    unsigned long long hi = (unsigned long long)w.z | ((unsigned long long)w.w << 32);
                    if (k < 8) lo &= ~0ull << (8 * k); else { lo = 0ull; hi &= ~0ull << (8 * (k - 8)); }
                    if (lo) u = base + ((__ffsll((long long)lo) - 1) >> 3);
                    else if (hi) u = base + 8 + ((__ffsll((long long)hi) - 1) >> 3);
                    else { u = base + 15; continue; }          // none of the 16: on to the next group
                    if (u >= N) break;
                }
                const int32_t L = s_L[u], pl = s_pl[u];
                const int2 memo = s_h[u];
                const int32_t mmL = (L >= 1 && L <= N) ? s_mm[L - 1] : 0x7fffffff;
                int32_t m, q;
                if (ver == 0) m = L;
                else if (memo.y == ver) m = L - memo.x;
                else { m = 1; q = pl; while (q >= 0 && s_pres[q]) { m++; q = s_pl[q]; } }
                if (L >= T.over_gap && L <= N && m >= mmL) {                      // :129
                    int32_t l2min = 0x7fffffff, l2max = -1, qlast = u;
                    q = u;
                    for (int32_t k = 0; k < m; k++) {
                        if (q < 0) break;                      // (only with a caller's broken order, reported by the caller)
                        const int32_t qn = s_pl[q], yv = s_y[q];   // (the next hop first: it is the loop's dependent chain)
                        s_pres[q] = 0;
                        T.mscr[cursor + k] = s_f[q];
                        l2min = yv < l2min ? yv : l2min;
                        l2max = yv > l2max ? yv : l2max;
                        qlast = q;
                        q = qn;
                    }
                    const int32_t la = s_x[u], lb = s_x[qlast];
                    ver++;
                    const int32_t t = s_t[u];
                    T.acc_m[t] = (uint32_t)m;
                    T.acc_L[t] = L;
                    T.acc_ms[t] = cursor;
                    T.acc_box[t] = make_int4(la < lb ? la : lb, la < lb ? lb : la, l2min, l2max);
                    cursor += (uint32_t)m;
                } else {
                    s_pres[u] = 0;                             // :143
                    if (m > 1) s_h[pl] = make_int2(L - m, ver);
                }
            }
        }
        __syncwarp();
        (void)rev;
    }
}
#endif

struct Stage {                               // bump allocator over the context's staging buffers
    uint8_t *base;
    size_t used;
    template <class T> T *take(int64_t count) {
        used = (used + 63) & ~(size_t)63;
        T *p = (T *)(base + used);
        used += (size_t)(count > 0 ? count : 1) * sizeof(T);
        return p;
    }
};

// Output staging for NBo blocks with NMo gene pairs: one device allocation and one page-locked host allocation of the same
// layout — the per-block arrays first, the per-gene-pair arrays behind them (they are filled and copied separately).
static int stage_layout(icl_ctx *c, int64_t NBo, int64_t NMo) {
    Status &st = c->st;
    // Order of the per-block arrays: what the collinearity-file writer reads first (ICL_FIELDS_WRITER copies less), and in
    // both groups the arrays the times pass does not touch before the ones it fills — those leave early, on the side stream
    // under that pass; only p_raw (and N, N1) are copied behind it:
    //   [pair score m member_off | p_raw] [dir L L1 L2 | N N1] [loc1 loc2 | member_idx]
    size_t members_at = 0, blocks_lean_end = 0, members_lean_end = 0, praw_at = 0, late_at = 0;
    auto layout = [&](Stage &g, icl_blocks &o) {
        o.pair = g.take<int32_t>(NBo); o.score = g.take<int64_t>(NBo); o.m = g.take<int32_t>(NBo);
        o.member_off = g.take<int64_t>(NBo + 1);
        o.p_raw = g.take<double>(NBo);
        praw_at = (size_t)((uint8_t *)o.p_raw - g.base);
        o.dir = g.take<int32_t>(NBo);
        blocks_lean_end = (size_t)((uint8_t *)o.dir - g.base);
        o.L = g.take<int32_t>(NBo); o.L1 = g.take<int32_t>(NBo); o.L2 = g.take<int32_t>(NBo);
        o.N = g.take<int32_t>(NBo);
        late_at = (size_t)((uint8_t *)o.N - g.base);
        o.N1 = g.take<int64_t>(NBo);
        o.loc1 = g.take<int32_t>(NMo);
        members_at = (size_t)((uint8_t *)o.loc1 - g.base);
        o.loc2 = g.take<int32_t>(NMo);
        o.member_idx = g.take<int32_t>(NMo);
        members_lean_end = (size_t)((uint8_t *)o.member_idx - g.base);
    };
    Stage ds{nullptr, 0};
    icl_blocks probe{};
    layout(ds, probe);                                         // size only
    const size_t bytes = ds.used + 64;
    if (c->d_stage.reserve(st, bytes)) return st.code;
    if (bytes > c->h_stage_cap) {
        if (c->h_stage) pin_free(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        if (pin_malloc((void **)&c->h_stage, bytes + bytes / 4) != cudaSuccess || !c->h_stage) {
            c->h_stage = nullptr;
            return fail(c, ICL_ERR_NOMEM, "icl_extract: out of page-locked host memory");
        }
        c->h_stage_cap = bytes + bytes / 4;
    }
    ds = Stage{c->d_stage.p, 0};
    Stage hs{c->h_stage, 0};
    memset(&c->dv, 0, sizeof(c->dv));
    memset(&c->hv, 0, sizeof(c->hv));
    layout(ds, c->dv);
    layout(hs, c->hv);
    c->stage_members_at = members_at;
    c->stage_bytes = ds.used;
    c->stage_blocks_lean_end = blocks_lean_end;
    c->stage_members_lean_end = members_lean_end;
    c->stage_praw_at = praw_at;
    c->stage_late_at = late_at;
    return 0;
}

// position e of the pair-major emission order -> rank position t (V space, direction-major): the 2 np positions of pair p
// are its forward rank positions followed by its anti-diagonal ones
ICL_HD int64_t emission_pos(int64_t e, int64_t n, const int32_t *ps, const int64_t *off) {
    const int32_t p = ps[e >> 1];
    const int64_t a = off[p], np_ = off[p + 1] - a, r = e - 2 * a;
    return r < np_ ? a + r : n + a + (r - np_);
}

static int phase_extract(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->present.reserve(st, V) || c->key_a.reserve(st, V) || c->key_b.reserve(st, V) || c->idx_a.reserve(st, V) ||
        c->tree_v.reserve(st, V) || c->acc_m.reserve(st, V + 2) || c->acc_L.reserve(st, V) || c->acc_ms.reserve(st, V) ||
        c->acc_box.reserve(st, V) || c->mscr.reserve(st, V) || c->seen.reserve(st, V) || c->bs_in.reserve(st, V + 2) ||
        c->bs.reserve(st, V + 2) || c->mo_in.reserve(st, V + 2) || c->mo.reserve(st, V + 2) || c->hint.reserve(st, V) || c->stats.reserve(st, 8) ||
        c->big_trees.reserve(st, V / ICL_TREE_SMALL + 2) || c->small_trees.reserve(st, V + 2))
        return st.code;
    ICL_CUDA(st, dev_memset(c->stats.p, 0, 8 * sizeof(int32_t), s));
    if (!c->order_on_device_used) {                            // (a device order initialises these in its key pass below)
        ICL_CUDA(st, dev_memset(c->acc_m.p, 0, (size_t)(V + 2) * sizeof(uint32_t), s));
        ICL_CUDA(st, dev_memset(c->hint.p, 0xFF, (size_t)V * sizeof(int2), s));
        ICL_CUDA(st, d2d(c->present.p, c->used.p, (size_t)V, s));  // drop usedtimes < 1 (:50 / :61)
    }
    const int2 *pd = c->pd.p;
    const int32_t *root = c->root.p, *rank2v = c->rank2v.p;
    const uint8_t *used = c->used.p;
    int32_t *stats = c->stats.p;
    // Rank positions are item-major and already in examination order, so a STABLE sort by tree root alone
    // groups every predecessor tree with its points in rank order.  Unused points sort behind everything.
    // The sort key is the root's position INSIDE its item (as many bits as the largest pair needs: 2 radix passes instead
    // of 3 on C2, 2 instead of 4 on C4): equal keys of different items stay item-major — the sort is stable — so every
    // (item, root) run is still contiguous and in rank order; the pass behind the sort turns the keys back into roots.
    // (Trees then come out ordered by position-inside-the-item first: neighbouring threads of the short-tree walk work on
    // different items and lose their shared cache lines — 93 -> 130 us on C2 — so the short key is used only where it
    // saves two passes or more.)
    int64_t maxnp = 1;
    for (int32_t p = 0; p < P; p++) maxnp = std::max<int64_t>(maxnp, c->h_pair_off[p + 1] - c->h_pair_off[p]);
    const bool rel = (bits_for(V) + 7) / 8 - (bits_for(maxnp) + 7) / 8 >= 2;
    const int vb = rel ? bits_for(maxnp) : bits_for(V);
    const uint32_t unused_key = rel ? (uint32_t)maxnp : (uint32_t)V;
    const int64_t *seg_off = c->d_seg_off.p;
    {
        uint32_t *key = (uint32_t *)c->key_a.p;                  // position inside the item < 2^24: 32-bit keys
        uint32_t *idx = c->idx_a.p;
        uint8_t *seen = c->seen.p;
        if (c->order_on_device_used) {
            // the order comes from the device's own sort: a permutation by construction, nothing to validate.  One pass:
            // the point examined at rank position t (undoing pandas' two reversals), its tree key, and the initial state
            // of the walk (present = used: :50 / :61 drop usedtimes < 1)
            const int32_t *itemof = c->itemof.p, *mmap = c->mmap.p, *ord = c->order_in.p;
            const int64_t *off = c->d_pair_off.p;
            int32_t *r2v = c->rank2v.p;
            uint32_t *acc_m = c->acc_m.p;
            int2 *hint = c->hint.p;
            uint8_t *present = c->present.p;
            ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t t) {
                const int32_t it = itemof[t];
                const bool rev = it >= P;
                const int64_t a = off[rev ? it - P : it], np_ = off[(rev ? it - P : it) + 1] - a;
                const int64_t r = t - (rev ? n : 0) - a;           // rank position inside the item
                const int32_t rf = ord[(rev ? n : 0) + a + (np_ - 1 - r)];
                const int64_t f = a + (np_ - 1 - rf);              // forward sorted position of the point
                const int32_t v = rev ? mmap[f] : (int32_t)f;
                r2v[t] = v;
                idx[t] = (uint32_t)t;
                key[t] = used[v] ? (uint32_t)(root[v] - (rel ? (rev ? n : 0) + a : 0)) : unused_key;
                acc_m[t] = 0u;
                hint[t] = make_int2(-1, -1);
                present[t] = used[t];
            }));
        } else {
            ICL_CUDA(st, dev_memset(c->seen.p, 0, (size_t)V, s));
            const int32_t *itemof = c->itemof.p;
            ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t t) {
                idx[t] = (uint32_t)t;
                const int32_t v = rank2v[t];
                if (v < 0 || v >= V || seen[v]) { atomic_or_i32(&stats[2], 1); key[t] = unused_key; return; }   // not a permutation
                seen[v] = 1;
                key[t] = used[v] ? (uint32_t)(root[v] - (rel ? seg_off[itemof[t]] : 0)) : unused_key;   // (v lies in the item of t by construction)
            }));
            // a duplicate may slip through the unsynchronised test above, but then some point is never seen
            ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t v) { if (!seen[v]) atomic_or_i32(&stats[2], 1); }));
        }
    }
    if (V > 0 && radix_sort_pairs32(st, c->tmp, (uint32_t *)c->key_a.p, (uint32_t *)c->key_b.p, c->idx_a.p, c->tree_v.p, V, vb, s)) return st.code;
    if (rel) {
        uint32_t *key = (uint32_t *)c->key_b.p;
        const uint32_t *tpos = c->tree_v.p;
        const int32_t *itemof = c->itemof.p;
        ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t j) {
            const uint32_t k = key[j];
            key[j] = k == unused_key ? (uint32_t)V : (uint32_t)(seg_off[itemof[tpos[j]]] + k);
        }));
    }
    {
        const uint32_t *key = (const uint32_t *)c->key_b.p;
        const uint32_t *tpos = c->tree_v.p;            // rank position of every sorted entry
        const int32_t *imap = c->imap.p, *l1 = c->loc1s.p, *l2 = c->loc2s.p;
        uint8_t *present = c->present.p;
        uint32_t *acc_m = c->acc_m.p, *acc_ms = c->acc_ms.p;
        int32_t *acc_L = c->acc_L.p, *mscr = c->mscr.p;
        int4 *acc_box = c->acc_box.p;
        int2 *hint = c->hint.p;
        const int32_t over_gap = c->prm.over_gap;
        const double cov = c->prm.coverage_ratio;
#ifndef ICL_EMU
        int2 *big = c->big_trees.p;
#endif
        TreeWalk W{n, V, key, tpos, rank2v, pd, imap, l1, l2, present, hint, acc_m, acc_ms, acc_L, mscr, acc_box, over_gap, cov};
#ifdef ICL_EMU
        ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t t0) {
            const uint32_t rt = key[t0];
            if (rt >= (uint32_t)V) return;                                      // unused tail
            if (t0 > 0 && key[t0 - 1] == rt) return;                            // not the head of its tree
            W.walk(t0, W.run_end(t0));
        }));
#else
        if (V > 0) {
            // tree heads -> two queues: trees of at most ICL_TREE_SMALL points (one thread each: compacted, so that every
            // lane of a warp has a tree to walk) and longer ones (one warp each, staged in shared memory)
            int2 *smallq = c->small_trees.p;
            const int64_t bigcap = V / ICL_TREE_SMALL + 2;
            k_tree_heads<<<148 * 8, 256, 0, s>>>(W, smallq, big, bigcap, stats);
            ICL_COUNT_KERNEL();
            // the two walks touch different trees: the staged one (few long trees, one warp each) runs beside the short
            // ones on the side stream
            ICL_CUDA(st, cudaEventRecord(c->ev_fork, s));
            ICL_CUDA(st, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
            ICL_CUDA(st, cudaStreamWaitEvent(c->stream3, c->ev_fork, 0));
            TreeArgs T{n, V, key, tpos, rank2v, pd, imap, l1, l2, big, bigcap, stats, acc_m, acc_ms, acc_L, mscr, acc_box, over_gap, cov};
            ICL_CUDA(st, cudaFuncSetAttribute(k_tree_walk_staged<ICL_TREE_CAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              ICL_TREE_SMEM(ICL_TREE_CAP)));
            k_tree_walk_staged<ICL_TREE_CAP, true><<<148 * 2, 32, ICL_TREE_SMEM(ICL_TREE_CAP), c->stream2>>>(T);
            ICL_COUNT_KERNEL();
            k_tree_walk_staged<ICL_TREE_HUGE, false><<<148 * 16, 32, ICL_TREE_SMEM(ICL_TREE_HUGE), c->stream3>>>(T);
            ICL_COUNT_KERNEL();
            ICL_CUDA(st, cudaEventRecord(c->ev_join3, c->stream3));
            ICL_CUDA(st, cudaEventRecord(c->ev_join, c->stream2));
            k_tree_walk_small<<<148 * 16, 128, 0, s>>>(W, smallq, stats);
            ICL_COUNT_KERNEL();
            ICL_CUDA(st, cudaStreamWaitEvent(s, c->ev_join, 0));
            ICL_CUDA(st, cudaStreamWaitEvent(s, c->ev_join3, 0));
            ICL_CUDA(st, cudaGetLastError());
        }
#endif
    }
    // Emission order (run_colliearity.py:118, collinearity.py:49-68): chromosome pair ascending, forward blocks before
    // anti-diagonal ones, acceptance order = rank order inside a direction.  Rank positions are direction-major (V space),
    // so the accepted positions are counted in the pair-major order t' -> t of emission_pos(): prefix sums number every
    // accepted block and every block the caller wants (m >= min_members; two 32-bit counts in one word) and the gene pairs
    // of those, all in emission order.
    {
        const uint32_t *acc_m = c->acc_m.p;
        const int32_t *ps = c->pair_s.p;
        const int64_t *off = c->d_pair_off.p;
        uint64_t *bs_in = c->bs_in.p;
        uint32_t *mo_in = c->mo_in.p;
        const uint32_t minm = (uint32_t)(c->min_members < 1 ? 1 : c->min_members);
        ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t e) {
            const uint32_t m = acc_m[emission_pos(e, n, ps, off)];
            const bool want = m >= minm;
            bs_in[e] = ((uint64_t)(m ? 1u : 0u) << 32) | (want ? 1u : 0u);
            mo_in[e] = want ? m : 0u;
        }));
    }
    if (exclusive_sum_u64(st, c->tmp, c->bs_in.p, c->bs.p, V, s)) return st.code;
    if (exclusive_sum_u32(st, c->tmp, c->mo_in.p, c->mo.p, V, s)) return st.code;
    // the counts go to the host on the side stream (page-locked words: [0] blocks << 32 | wanted, [1] gene pairs, [2..3]
    // stats, [4] rank hand-back flag); extract_counts() waits for them
    c->h_small[4] = 0;
    cudaStream_t s2 = c->stream2 ? c->stream2 : s;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventRecord(c->ev_fork, s));
    ICL_CUDA(st, cudaStreamWaitEvent(s2, c->ev_fork, 0));
#endif
    ICL_CUDA(st, d2h(&c->h_small[0], c->bs.p + V, sizeof(uint64_t), s2));
    ICL_CUDA(st, d2h(&c->h_small[1], c->mo.p + V, sizeof(uint32_t), s2));
    ICL_CUDA(st, d2h(&c->h_small[2], c->stats.p, 4 * sizeof(int32_t), s2));
    if (c->rank_flag_pending && c->order_on_device_used)       // hand-back flag icl_rank_device left unread (no sync there)
        ICL_CUDA(st, d2h(&c->h_small[4], c->host_flag.p + c->rank_flag_slot, sizeof(int32_t), s2));
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventRecord(c->ev_join, s2));
#endif
    return 0;
}

// Wait for the counts of phase_extract: accepted blocks, blocks / gene pairs the caller gets; errors the walk found.
static int extract_counts(icl_ctx *c) {
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventSynchronize(c->ev_join));
#endif
    const uint64_t tot_bs = c->h_small[0];
    const uint32_t tot_m = *(const uint32_t *)&c->h_small[1];
    const int32_t *hs = (const int32_t *)&c->h_small[2];
    const int32_t rank_any = *(const int32_t *)&c->h_small[4];
    c->rank_flag_pending = false;
    if (rank_any) return fail(c, ICL_ERR_STATE, "icl_rank_device: a segment needs the host rank (std::sort depth limit)");
    if (hs[2]) return fail(c, ICL_ERR_ARG, "icl_extract: argsort array is not a permutation of its pair segment");
    c->NB = (int64_t)(tot_bs >> 32);
    c->NBo = (int64_t)(uint32_t)tot_bs;
    c->NMo = tot_m;
    return 0;
}

// One thread per accepted block (emission order): the row slice of its bounding box for the times / p-value pass, and the
// emission position of every block the caller gets.  `cap` blocks fit the arrays: the exact count when the host already
// has it, else the bound of extract_block_bound() — the pass then runs (and the p-value kernel behind it) while the
// counts are still on their way to the host.
static int extract_block_slices(icl_ctx *c, int64_t cap) {
    Status &st = c->st;
    const int64_t n = c->n, V = 2 * n;
    cudaStream_t s = c->stream;
    if (c->b_j0.reserve(st, cap) || c->b_j1.reserve(st, cap) || c->b_l2min.reserve(st, cap) || c->b_l2max.reserve(st, cap) ||
        c->b_N.reserve(st, cap) || c->b_N1.reserve(st, cap) || c->o_e.reserve(st, cap + 1))
        return st.code;
    const uint32_t *acc_m = c->acc_m.p;
    const uint64_t *bs = c->bs.p;
    const int32_t *rowptr = c->rowptr.p, *ps = c->pair_s.p;
    const int4 *acc_box = c->acc_box.p;
    const int32_t *rmin = c->d_rmin.p;
    const int64_t *rb_ = c->d_rb.p, *off = c->d_pair_off.p;
    int32_t *b_j0 = c->b_j0.p, *b_j1 = c->b_j1.p, *b_l2min = c->b_l2min.p, *b_l2max = c->b_l2max.p, *b_N = c->b_N.p;
    int64_t *b_N1 = c->b_N1.p;
    uint32_t *o_e = c->o_e.p;
    const uint32_t minm = (uint32_t)(c->min_members < 1 ? 1 : c->min_members);
    ICL_CUDA(st, foreach(s, V, ICL_LAMBDA(int64_t e) {
        const int32_t p = ps[e >> 1];
        const int64_t t = emission_pos(e, n, ps, off);
        const uint32_t m = acc_m[t];
        if (!m) return;
        const uint64_t at = bs[e];
        const uint32_t b = (uint32_t)(at >> 32);
        if ((int64_t)b >= cap) return;                                      // (cannot happen: cap is the count or its bound)
        const int4 box = acc_box[t];
        b_N[b] = 0;                                                         // (the p-value pass adds into these)
        b_N1[b] = 0;
        b_l2min[b] = box.z;
        b_l2max[b] = box.w;
        b_j0[b] = rowptr[rb_[p] + (box.x - rmin[p])];
        b_j1[b] = rowptr[rb_[p] + (box.y - rmin[p]) + 1];
        if (m >= minm) o_e[(uint32_t)at] = (uint32_t)e;
    }));
    return 0;
}

// No comparison accepts more blocks than this: a block takes m present points with m / L > coverage_ratio and
// L >= over_gap (collinearity.py:129), i.e. at least the smallest m that passes for L = over_gap, and a point is taken once.
static int64_t extract_block_bound(const icl_ctx *c) {
    const int64_t V = 2 * c->n;
    const int32_t L = c->prm.over_gap < 1 ? 1 : c->prm.over_gap;
    int32_t mmin = 1;
    while (mmin < L && !((double)mmin / (double)L > c->prm.coverage_ratio)) mmin++;
    return V / mmin + 1;
}

// The output records of the blocks the caller gets (everything but N, N1 and the p value, which follow the times pass),
// one thread each, on stream `s`.
static int extract_output_records(icl_ctx *c, cudaStream_t s) {
    Status &st = c->st;
    const int64_t n = c->n, NBo = c->NBo, NMo = c->NMo;
    if (c->o_src.reserve(st, NBo + 1) || c->o_ms.reserve(st, NBo + 1)) return st.code;
    const uint32_t *acc_m = c->acc_m.p, *acc_ms = c->acc_ms.p, *o_e = c->o_e.p, *mo = c->mo.p;
    const uint64_t *bs = c->bs.p;
    const int32_t *acc_L = c->acc_L.p, *S = c->S.p, *ps = c->pair_s.p, *rank2v = c->rank2v.p;
    const int4 *acc_box = c->acc_box.p;
    const int64_t *off = c->d_pair_off.p;
    uint32_t *o_src = c->o_src.p, *o_ms = c->o_ms.p;
    const icl_blocks o = c->dv;
    ICL_CUDA(st, foreach(s, NBo, ICL_LAMBDA(int64_t g) {
        if (g == 0) o.member_off[NBo] = NMo;
        const int64_t e = o_e[g];
        const int64_t t = emission_pos(e, n, ps, off);
        const int4 box = acc_box[t];
        o_src[g] = (uint32_t)(bs[e] >> 32);
        o_ms[g] = acc_ms[t];
        o.pair[g] = ps[e >> 1];
        o.dir[g] = t >= n ? 1 : 0;
        o.score[g] = S[rank2v[t]];
        o.m[g] = (int32_t)acc_m[t];
        o.L[g] = acc_L[t];
        o.L1[g] = box.y - box.x + 1;
        o.L2[g] = box.w - box.z + 1;
        o.member_off[g] = mo[e];
    }));
    if (NBo == 0) ICL_CUDA(st, dev_memset(c->dv.member_off, 0, sizeof(int64_t), s));
    return 0;
}

// The gene pairs of the output blocks, one thread each (binary search for its block): the walk listed a block's members
// from the top towards the root — loc1 descending forward, ascending on the anti-diagonal — and the output wants loc1
// ascending (:140).  Independent of the times / p-value pass: issued on the side stream next to it, its copy behind it.
static int phase_members(icl_ctx *c, cudaStream_t s) {
    Status &st = c->st;
    const int64_t NBo = c->NBo, NMo = c->NMo;
    const uint32_t *o_ms = c->o_ms.p, *perm = c->perm.p;
    const int32_t *mscr = c->mscr.p, *l1 = c->loc1s.p, *l2 = c->loc2s.p;
    const icl_blocks o = c->dv;
    ICL_CUDA(st, foreach(s, NMo, ICL_LAMBDA(int64_t t) {
        int64_t lo = 0, hi = NBo - 1;                           // last block with member_off[g] <= t
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo + 1) >> 1);
            if (o.member_off[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int64_t k = t - o.member_off[lo], m = o.member_off[lo + 1] - o.member_off[lo];
        const int32_t i = mscr[o_ms[lo] + (uint32_t)(o.dir[lo] ? k : m - 1 - k)];
        o.loc1[t] = l1[i];
        o.loc2[t] = l2[i];
        o.member_idx[t] = (int32_t)perm[i];
    }));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// phase: times / p value (collinearity.py:131-139, 146-153).  `times` is shared by the two directions of
// a chromosome pair and by all of its blocks in acceptance order, so one warp owns a pair and walks its
// blocks (forward then anti-diagonal); the box scan over the contiguous row slice is lane-parallel and
// `times` lives in shared memory (uint16) when the pair fits.
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

struct PvalArgs {
    int64_t n;
    int32_t P, g0, smem_cap;
    const int32_t *pair_order, *l2, *b_j0, *b_j1, *b_l2min, *b_l2max;
    const int64_t *off;
    const uint64_t *bs;              // high word: accepted blocks before an emission position (pair p: forward blocks from 2 off[p], anti-diagonal from 2 off[p] + np)
    ICL_HD uint32_t blocks_before(int64_t e) const { return (uint32_t)(bs[e] >> 32); }
    int32_t *times, *b_N;
    int64_t *b_N1;
};

#ifndef ICL_EMU
// MODE 0: times (uint16) and loc2 of the pair in shared memory; 1: times only; 2: everything in global memory
template <int MODE>
__device__ __forceinline__ void pvalue_pair(const PvalArgs &A, int32_t p, uint16_t *tsm, int32_t *l2sm) {
    const int lane = threadIdx.x;
    const int64_t a = A.off[p], np_ = A.off[p + 1] - a;
    if (MODE <= 1) {
        for (int64_t i = lane; i < np_; i += 32) tsm[i] = 1;                     // get_matrix :34
        if (MODE == 0)
            for (int64_t i = lane; i < np_; i += 32) l2sm[i] = A.l2[a + i];
    } else {
        for (int64_t i = lane; i < np_; i += 32) A.times[a + i] = 1;
    }
    __syncwarp();
    volatile int32_t *tg = A.times;
    for (int d = 0; d < 2; d++) {
        const uint32_t b0 = A.blocks_before(2 * a + d * np_), b1 = A.blocks_before(2 * a + (d + 1) * np_);
        // block records are fetched 32 at a time (one per lane, coalesced) and handed out by shuffle
        int32_t rj0 = 0, rj1 = 0, rlo = 0, rhi = 0;
        if (b0 + lane < b1) { rj0 = A.b_j0[b0 + lane]; rj1 = A.b_j1[b0 + lane]; rlo = A.b_l2min[b0 + lane]; rhi = A.b_l2max[b0 + lane]; }
        for (uint32_t base = b0; base < b1; base += 32) {
            const int32_t cj0 = rj0, cj1 = rj1, clo = rlo, chi = rhi;
            const uint32_t nb = base + 32 + lane;
            if (nb < b1) { rj0 = A.b_j0[nb]; rj1 = A.b_j1[nb]; rlo = A.b_l2min[nb]; rhi = A.b_l2max[nb]; }
            const uint32_t cnt = b1 - base < 32 ? b1 - base : 32;
            int32_t myN = 0;
            unsigned long long myN1 = 0;
            for (uint32_t i = 0; i < cnt; i++) {
                const int32_t j0 = __shfl_sync(0xffffffffu, cj0, i), j1 = __shfl_sync(0xffffffffu, cj1, i);
                const int32_t lo = __shfl_sync(0xffffffffu, clo, i), hi = __shfl_sync(0xffffffffu, chi, i);
                uint32_t N = 0;
                unsigned long long N1 = 0;
                for (int32_t q = j0 + lane; q < j1; q += 32) {                    // :134-137 inclusive box
                    const int32_t y = MODE == 0 ? l2sm[q - a] : __ldg(A.l2 + q);
                    if (y >= lo && y <= hi) {
                        uint32_t t;
                        if (MODE <= 1) { t = tsm[q - a]; tsm[q - a] = (uint16_t)(t + 1); }
                        else { t = (uint32_t)tg[q]; tg[q] = (int32_t)t + 1; }
                        N1 += t;                                                  // :148
                        N++;                                                      // :149, times += 1 :150
                    }
                }
                N = __reduce_add_sync(0xffffffffu, N);
                if (MODE <= 1) {                                  // uint16 counters over <= 102 400 points: the sum fits 32 bits
                    N1 = __reduce_add_sync(0xffffffffu, (uint32_t)N1);
                } else {
                    for (int o = 16; o > 0; o >>= 1) N1 += __shfl_xor_sync(0xffffffffu, N1, o);
                }
                __syncwarp();                                                     // times of this block visible to the next
                if (lane == (int)i) { myN = (int32_t)N; myN1 = N1; }
            }
            if ((uint32_t)lane < cnt) { A.b_N[base + lane] = myN; A.b_N1[base + lane] = (int64_t)myN1; }
        }
    }
}

__global__ void __launch_bounds__(32) k_pvalue(PvalArgs A, int32_t l2_cap) {
    extern __shared__ __align__(16) unsigned char psm[];
    const int32_t p = A.pair_order[blockIdx.x];
    const int64_t np_ = A.off[p + 1] - A.off[p];
    const uint32_t nb = A.blocks_before(2 * A.off[p + 1]) - A.blocks_before(2 * A.off[p]);
    if (nb == 0) return;
    uint16_t *tsm = (uint16_t *)psm;
    int32_t *l2sm = (int32_t *)(psm + (((size_t)A.smem_cap * 2 + 15) & ~(size_t)15));
    // uint16 counters: a point is covered by fewer boxes than the pair has blocks
    if (np_ <= l2_cap && nb < 65000u) pvalue_pair<0>(A, p, tsm, l2sm);
    else if (np_ <= A.smem_cap && nb < 65000u) pvalue_pair<1>(A, p, tsm, l2sm);
    else pvalue_pair<2>(A, p, tsm, l2sm);
}
#endif

#ifndef ICL_EMU
// times / N / N1 with several warps per chromosome pair.  `times` couples the blocks of a pair only through the points
// they share, so the POINTS are divided instead of the blocks: every warp owns a contiguous share of the pair's points,
// walks ALL blocks of the pair in acceptance order (forward, then anti-diagonal) and treats the part of each box's row
// slice that falls into its share — the counters of a point are only ever touched by its owner, in block order, so the
// warps never wait for each other; the per-block sums meet in b_N / b_N1 through atomic adds (zeroed before the launch).
// Blocks that miss a warp's share cost it one ballot per 32 blocks.  (The one-warp-per-pair kernel was a chain of ~900
// dependent block steps on the heaviest pair: 0.40 ms on C2 with 2.4 warps per SM.)
// TS: counters and loc2 of the pair in shared memory (pairs up to l2_cap points); else counters in global memory.
template <bool TS>
__global__ void __launch_bounds__(1024) k_pvalue_owned(PvalArgs A) {
    extern __shared__ __align__(16) unsigned char psm[];
    const int32_t p = A.pair_order[blockIdx.x];
    const int64_t a = A.off[p], np_ = A.off[p + 1] - a;
    const uint32_t bf0 = A.blocks_before(2 * a), bf1 = A.blocks_before(2 * a + np_), br0 = bf1, br1 = A.blocks_before(2 * a + 2 * np_);
    if (bf1 - bf0 + br1 - br0 == 0) return;
    uint32_t *tsm = (uint32_t *)psm;                           // TS: one counter per point of the pair
    int32_t *l2sm = (int32_t *)(psm + (((size_t)np_ * 4 + 15) & ~(size_t)15));
    if (TS) {
        for (int64_t i = threadIdx.x; i < np_; i += blockDim.x) { tsm[i] = 1; l2sm[i] = A.l2[a + i]; }   // get_matrix :34
    } else {
        for (int64_t i = threadIdx.x; i < np_; i += blockDim.x) A.times[a + i] = 1;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int32_t own0 = (int32_t)(a + np_ * w / nw), own1 = (int32_t)(a + np_ * (w + 1) / nw);
    const uint32_t FULL = 0xffffffffu;
    for (int d = 0; d < 2; d++) {
        const uint32_t b0 = d ? br0 : bf0, b1 = d ? br1 : bf1;
        for (uint32_t base = b0; base < b1; base += 32) {
            const uint32_t b = base + lane;
            int32_t rj0 = 0, rj1 = 0, rlo = 0, rhi = 0;
            if (b < b1) { rj0 = A.b_j0[b]; rj1 = A.b_j1[b]; rlo = A.b_l2min[b]; rhi = A.b_l2max[b]; }
            rj0 = rj0 > own0 ? rj0 : own0;                     // the slice inside my share
            rj1 = rj1 < own1 ? rj1 : own1;
            uint32_t todo = __ballot_sync(FULL, b < b1 && rj0 < rj1);
            while (todo) {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                const int32_t j0 = __shfl_sync(FULL, rj0, i), j1 = __shfl_sync(FULL, rj1, i);
                const int32_t lo = __shfl_sync(FULL, rlo, i), hi = __shfl_sync(FULL, rhi, i);
                uint32_t N = 0;
                unsigned long long N1 = 0;
                for (int32_t q = j0 + lane; q < j1; q += 32) {                    // :134-137 inclusive box
                    const int32_t y = TS ? l2sm[q - a] : __ldg(A.l2 + q);
                    if (y >= lo && y <= hi) {
                        uint32_t t;
                        if (TS) { t = tsm[q - a]; tsm[q - a] = t + 1; }
                        else { t = (uint32_t)A.times[q]; A.times[q] = (int32_t)t + 1; }
                        N1 += t;                                                  // :148
                        N++;                                                      // :149, times += 1 :150
                    }
                }
                N = __reduce_add_sync(FULL, N);
                for (int o = 16; o > 0; o >>= 1) N1 += __shfl_xor_sync(FULL, N1, o);
                __syncwarp();                                  // my lanes' counters before the next block of my share
                if (lane == 0 && N) {
                    atomicAdd(&A.b_N[base + i], (int32_t)N);
                    atomicAdd((unsigned long long *)&A.b_N1[base + i], N1);
                }
            }
        }
    }
}
#endif

// Order-free form of the `times` accounting: times[q] before block b = 1 + number of blocks accepted earlier in
// the same pair whose box contains q, so  N1(b) = N(b) + sum over the points q of box(b) of rank_q(b),  where
// rank_q(b) counts the earlier blocks containing q.  Enumerate the (point, block) incidences of all boxes in
// parallel, radix-sort them by (point, block index) — block indices grow in acceptance order inside a pair,
// forward blocks before anti-diagonal ones — and the rank is the position inside the point's run.
// Returns handled = false when the incidence list would be unreasonably large (then the sequential kernel runs).
static int phase_pvalue_incidence(icl_ctx *c, bool *handled) {
    Status &st = c->st;
    const int64_t NB = c->NB;
    cudaStream_t s = c->stream;
    *handled = false;
    if (NB == 0) { *handled = true; return 0; }
    if (c->inc_cnt.reserve(st, NB + 2) || c->inc_off.reserve(st, NB + 2)) return st.code;
    const int32_t *l2 = c->loc2s.p, *b_j0 = c->b_j0.p, *b_j1 = c->b_j1.p, *b_lo = c->b_l2min.p, *b_hi = c->b_l2max.p;
    int32_t *b_N = c->b_N.p;
    int64_t *b_N1 = c->b_N1.p;
    {
        uint32_t *cnt = c->inc_cnt.p;
        ICL_CUDA(st, foreach(s, NB, ICL_LAMBDA(int64_t b) {
            const int32_t lo = b_lo[b], hi = b_hi[b];
            int32_t N = 0;
            for (int32_t q = b_j0[b]; q < b_j1[b]; q++) {                  // :134-137 inclusive box over the row slice
                const int32_t y = l2[q];
                N += (y >= lo && y <= hi);
            }
            b_N[b] = N;                                                     // :149
            cnt[b] = (uint32_t)N;
            b_N1[b] = N;                                                    // every point of the box counts once (times = 1)
        }));
    }
    if (exclusive_sum_u32(st, c->tmp, c->inc_cnt.p, c->inc_off.p, NB, s)) return st.code;
    // the 32-bit scan wraps when the boxes of all blocks hold 2^32 incidences or more; a box holds fewer than 2^24 points
    // (pair size limit), so a wrap always shows as a DEcrease between neighbours — checked before the total is trusted
    ICL_CUDA(st, dev_memset(c->stats.p + 1, 0, sizeof(int32_t), s));
    {
        const uint32_t *ioff = c->inc_off.p;
        int32_t *stats = c->stats.p;
        ICL_CUDA(st, foreach(s, NB, ICL_LAMBDA(int64_t b) { if (ioff[b + 1] < ioff[b]) atomic_or_i32(&stats[1], 1); }));
    }
    uint32_t total = 0;
    int32_t wrapped = 0;
    ICL_CUDA(st, d2h(&total, c->inc_off.p + NB, sizeof(uint32_t), s));
    ICL_CUDA(st, d2h(&wrapped, c->stats.p + 1, sizeof(int32_t), s));
    ICL_CUDA(st, dev_sync(s));
    const int64_t I = total;
    if (wrapped || I > (1ll << 28)) return 0;                               // pathological nesting: sequential kernel instead
    *handled = true;
    if (I == 0) return 0;
    if (c->key_a.reserve(st, I) || c->key_b.reserve(st, I) || c->inc_start.reserve(st, I + 2)) return st.code;
    const int sb = bits_for(NB);
    {
        const uint32_t *ioff = c->inc_off.p;
        uint64_t *key = c->key_a.p;
        ICL_CUDA(st, foreach(s, NB, ICL_LAMBDA(int64_t b) {
            const int32_t lo = b_lo[b], hi = b_hi[b];
            uint32_t e = ioff[b];
            for (int32_t q = b_j0[b]; q < b_j1[b]; q++) {
                const int32_t y = l2[q];
                if (y >= lo && y <= hi) key[e++] = ((uint64_t)(uint32_t)q << sb) | (uint64_t)b;
            }
        }));
    }
    if (radix_sort_keys(st, c->tmp, c->key_a.p, c->key_b.p, I, sb + bits_for(c->n), s)) return st.code;
    {
        const uint64_t *key = c->key_b.p;
        int32_t *start = c->inc_start.p;
        ICL_CUDA(st, foreach(s, I, ICL_LAMBDA(int64_t j) {
            start[j] = (j == 0 || (key[j] >> sb) != (key[j - 1] >> sb)) ? (int32_t)j : 0;
        }));
        if (inclusive_max_i32(st, c->tmp, c->inc_start.p, I, s)) return st.code;  // first incidence of the point's run
        const uint64_t bmask = (1ull << sb) - 1;
        ICL_CUDA(st, foreach(s, I, ICL_LAMBDA(int64_t j) {
            const int64_t rank = j - start[j];                              // earlier blocks of the pair containing the point
            if (rank) atomic_add_i64(&b_N1[key[j] & bmask], rank);          // :148 with :150 of the earlier blocks
        }));
    }
    return 0;
}

static int phase_pvalue(icl_ctx *c) {
    Status &st = c->st;
    const int64_t n = c->n, NB = c->NB;
    const int32_t P = c->P;
    cudaStream_t s = c->stream;
    if (c->times.reserve(st, n) || c->b_N.reserve(st, NB) || c->b_N1.reserve(st, NB)) return st.code;
    PvalArgs A;
    A.n = n; A.P = P; A.g0 = c->prm.g0;
    A.pair_order = c->d_pair_order.p; A.l2 = c->loc2s.p;
    A.b_j0 = c->b_j0.p; A.b_j1 = c->b_j1.p; A.b_l2min = c->b_l2min.p; A.b_l2max = c->b_l2max.p;
    A.off = c->d_pair_off.p; A.bs = c->bs.p; A.times = c->times.p; A.b_N = c->b_N.p; A.b_N1 = c->b_N1.p;
    if (P == 0) return 0;
    // Small pairs (times and loc2 of the pair fit in shared memory, few blocks per pair) are fastest on the
    // sequential warp-per-pair kernel; large pairs serialise there and take the order-free incidence path.
    int64_t maxnp_all = 1;
    for (int32_t p = 0; p < P; p++) maxnp_all = std::max<int64_t>(maxnp_all, c->h_pair_off[p + 1] - c->h_pair_off[p]);
    const bool prefer_seq = c->pvalue_mode != 2;               // the incidence form stays selectable (env ICL_B200_PVALUE=inc)
    bool handled = false;
    if (!prefer_seq && phase_pvalue_incidence(c, &handled)) return st.code;
    if (!handled) {
#ifndef ICL_EMU
    int64_t maxnp = 1;
    for (int32_t p = 0; p < P; p++) maxnp = std::max<int64_t>(maxnp, c->h_pair_off[p + 1] - c->h_pair_off[p]);
    if (c->pvalue_mode == 3) {                                 // the one-warp-per-pair kernel (A/B, env ICL_B200_PVALUE=warp)
        // shared memory per CTA: uint16 times for up to smem_cap points (+ int32 loc2 for up to l2_cap points)
        A.smem_cap = (int32_t)std::min<int64_t>(maxnp, 100 * 1024);
        int32_t l2_cap = (int32_t)std::min<int64_t>(maxnp, 12 * 1024);              // keeps >= 3 CTAs per SM
        size_t smem = (((size_t)A.smem_cap * 2 + 15) & ~(size_t)15) + (size_t)l2_cap * 4;
        ICL_CUDA(st, cudaFuncSetAttribute(k_pvalue, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        k_pvalue<<<(unsigned)P, 32, smem, s>>>(A, l2_cap);
    } else {
        // warps per pair: enough that a warp's share of the blocks is short, few enough that several pairs fit an SM
        const bool ts = maxnp <= 27 * 1024;                    // counters + loc2 of the pair: 8 B per point in shared memory
        static const int pv_threads = [] { const char *e = getenv("ICL_B200_PV_THREADS"); return e ? atoi(e) : 0; }();
        const int threads = pv_threads >= 64 && pv_threads <= 1024 ? (pv_threads & ~31)
                            : maxnp <= 1024 ? 128 : maxnp <= 4096 ? 512 : 1024;   // (sweep 128 / 256 / 512 / 1024 threads, ms: C2
                                                                                    // 0.23 / 0.14 / 0.096 / 0.091, C4 0.84 / 0.46 / 0.28 / 0.24, C3 8.1 / 4.0 / 2.2 / 1.1)
        if (ts) {
            const size_t smem = (((size_t)maxnp * 4 + 15) & ~(size_t)15) + (size_t)maxnp * 4;
            ICL_CUDA(st, cudaFuncSetAttribute(k_pvalue_owned<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            k_pvalue_owned<true><<<(unsigned)P, threads, smem, s>>>(A);
        } else {
            k_pvalue_owned<false><<<(unsigned)P, threads, 0, s>>>(A);
        }
    }
    ICL_COUNT_KERNEL();
    ICL_CUDA(st, cudaGetLastError());
#else
    A.smem_cap = 0;
    for (int32_t jp = 0; jp < P; jp++) {
        int32_t p = A.pair_order[jp];
        int64_t a = A.off[p], np_ = A.off[p + 1] - a;
        for (int64_t i = 0; i < np_; i++) A.times[a + i] = 1;
        for (int d = 0; d < 2; d++) {
            uint32_t b0 = A.blocks_before(2 * a + d * np_), b1 = A.blocks_before(2 * a + (d + 1) * np_);
            for (uint32_t b = b0; b < b1; b++) {
                int32_t N = 0, lo = A.b_l2min[b], hi = A.b_l2max[b];
                int64_t N1 = 0;
                for (int32_t q = A.b_j0[b]; q < A.b_j1[b]; q++)
                    if (A.l2[q] >= lo && A.l2[q] <= hi) { N1 += A.times[q]; N++; A.times[q]++; }
                A.b_N[b] = N;
                A.b_N1[b] = N1;
            }
        }
    }
#endif
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: icl_extract
// ------------------------------------------------------------------------------------------------
extern "C" void icl_free_blocks(icl_blocks *b) {
    if (b) memset(b, 0, sizeof(*b));        // the arrays live in page-locked staging owned by the context
}


// ------------------------------------------------------------------------------------------------
// numpy's SCALAR argsort, reproduced exactly on the device.
//
// numpy.argsort(kind='quicksort') on int64 runs npysort/quicksort.cpp `aquicksort_` wherever no SIMD sort is dispatched
// (CPUs without AVX2 / AVX-512, or NPY_DISABLE_CPU_FEATURES as used for the fixtures under tests/golden/):
//   depth limit cdepth = 2 * floor(log2(n)); ranges longer than 16 are partitioned — median of three on
//   (first, middle, last) with index swaps, pivot parked at last - 1, Hoare scans `do ++pi while v[pi] < vp`,
//   `do --pj while vp < v[pj]`, swap until they cross, pivot to pi, larger side pushed, smaller side continued, both at
//   cdepth - 1; a range popped with cdepth < 0 is heap sorted; ranges of at most 16 elements are insertion sorted.
// (Restated from the numpy source and checked against numpy 2.3.5 with its SIMD sorts disabled on 2 330 arrays.)
// Its tie permutation is a function of the keys alone, and every step can be done in parallel WITHOUT changing it:
//  * Hoare partition: the k-th element from the left that is not below the pivot is swapped with the k-th element from
//    the right that is not above it, for as long as the former lies left of the latter — the stops are found on the
//    untouched array because a swap never moves anything the scans have yet to pass.  One warp lists both kinds of
//    stops with ballots, counts the crossing point and performs the swaps at once;
//  * sub-ranges are independent (only their depth travels with them), so the order in which they are taken is free;
//  * insertion sort is a stable sort of at most 16 elements: 32 small ranges are sorted at a time, one per lane.
// One warp per (pair, direction) segment; keys int32 (read only), indices in global memory (L2 resident).
// ------------------------------------------------------------------------------------------------
#ifndef ICL_EMU
struct NpSortArgs {
    const int32_t *key;        // [2n] segments back to back: the REVERSED score arrays (what pandas hands to numpy)
    int32_t *idx;              // [2n] out: argsort of every segment (segment-local indices)
    int32_t *scr;              // [4n] scratch: per segment 2 x len stop lists
    const int64_t *off;        // [segments + 1] start of every segment in key / idx
    const int32_t *order;      // segment order (longest first)
    int32_t smem_cap;          // segments of at most smem_cap elements keep keys and indices in shared memory
};

#define ICL_NPS_SMALL 15         // SMALL_QUICKSORT of this numpy build: ranges with pr - pl <= 15 are insertion sorted
#define ICL_NPS_STACK 160

template <class IdxT>
__device__ __forceinline__ void nps_heapsort(const int32_t *v, IdxT *tosort, int32_t n) {   // aheapsort_, one lane
    IdxT *a = tosort - 1;
    int32_t i, j, l, tmp;
    for (l = n >> 1; l > 0; --l) {
        tmp = a[l];
        for (i = l, j = l << 1; j <= n;) {
            if (j < n && v[a[j]] < v[a[j + 1]]) j += 1;
            if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
        }
        a[i] = tmp;
    }
    for (; n > 1;) {
        tmp = a[n]; a[n] = a[1]; n -= 1;
        for (i = 1, j = 2; j <= n;) {
            if (j < n && v[a[j]] < v[a[j + 1]]) j++;
            if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
        }
        a[i] = tmp;
    }
}

// aquicksort_ on x[pl..pr] with depth budget cdepth, by ONE lane (sub-ranges of at most ICL_NPS_LANE elements: below
// that size a warp-wide partition costs more in fixed overhead than a lane needs for the whole sub-range)
#define ICL_NPS_LANE 96
// `popped`: the range comes off the stack (numpy checks the depth budget only then, not for the side it continues with)
template <class IdxT>
__device__ void nps_seq(const int32_t *v, IdxT *x, int32_t pl, int32_t pr, int cdepth, bool popped) {
    int32_t slo[10], shi[10];
    int sd[10], sp = 0;
    for (;;) {
        if (popped && cdepth < 0) {
            nps_heapsort(v, x + pl, pr - pl + 1);
        } else {
            while (pr - pl > ICL_NPS_SMALL) {
                const int32_t pm = pl + ((pr - pl) >> 1);
                int32_t t;
                if (v[x[pm]] < v[x[pl]]) { t = x[pm]; x[pm] = x[pl]; x[pl] = t; }
                if (v[x[pr]] < v[x[pm]]) { t = x[pr]; x[pr] = x[pm]; x[pm] = t; }
                if (v[x[pm]] < v[x[pl]]) { t = x[pm]; x[pm] = x[pl]; x[pl] = t; }
                const int32_t vp = v[x[pm]];
                int32_t pi = pl, pj = pr - 1;
                t = x[pm]; x[pm] = x[pj]; x[pj] = t;
                for (;;) {
                    do { ++pi; } while (v[x[pi]] < vp);
                    do { --pj; } while (vp < v[x[pj]]);
                    if (pi >= pj) break;
                    t = x[pi]; x[pi] = x[pj]; x[pj] = t;
                }
                t = x[pi]; x[pi] = x[pr - 1]; x[pr - 1] = t;
                if (pi - pl < pr - pi) { slo[sp] = pi + 1; shi[sp] = pr; pr = pi - 1; }
                else { slo[sp] = pl; shi[sp] = pi - 1; pl = pi + 1; }
                sd[sp++] = --cdepth;
            }
            for (int32_t pi = pl + 1; pi <= pr; ++pi) {                  // insertion sort
                const int32_t vi = x[pi], vp = v[vi];
                int32_t pj = pi;
                while (pj > pl && vp < v[x[pj - 1]]) { x[pj] = x[pj - 1]; pj--; }
                x[pj] = vi;
            }
        }
        if (sp == 0) break;
        sp--;
        pl = slo[sp]; pr = shi[sp]; cdepth = sd[sp];
        popped = true;
    }
}

#define ICL_NPS_WARPS 8
// One CTA of ICL_NPS_WARPS warps per segment.  The warps share a stack of sub-ranges (any order is fine: sub-ranges are
// disjoint and carry their own depth budget): a warp pops a range, partitions it, pushes the larger side and goes on
// with the smaller one until what is left fits a lane; those leftovers are sorted 32 at a time, one per lane.
template <class IdxT>
__global__ void __launch_bounds__(32 * ICL_NPS_WARPS) k_np_argsort(NpSortArgs A) {
    __shared__ int32_t st_lo[ICL_NPS_STACK], st_hi[ICL_NPS_STACK], st_d[ICL_NPS_STACK];
    __shared__ int32_t sm_lo[ICL_NPS_WARPS][32], sm_hi[ICL_NPS_WARPS][32], sm_d[ICL_NPS_WARPS][32], sm_p[ICL_NPS_WARPS][32];
    __shared__ int32_t s_sp, s_pending, s_lock;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t FULL = 0xffffffffu;
    const int32_t it = A.order[blockIdx.x];
    const int64_t base = A.off[it];
    const int32_t len = (int32_t)(A.off[it + 1] - base);
    if (len <= 0) return;
    extern __shared__ int32_t nps_sm[];
    // IdxT = uint16_t: every segment fits shared memory with 16-bit indices (6 B per element, more CTAs per SM);
    // IdxT = int32_t: segments beyond smem_cap work on global memory
    const bool fits = sizeof(IdxT) == 2 || len <= A.smem_cap;
    const int32_t *v = A.key + base;
    IdxT *x;
    if (fits) {                                              // the whole segment on chip: every gather is a shared-memory load
        int32_t *vs = nps_sm;
        IdxT *xs = (IdxT *)(nps_sm + A.smem_cap);
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) { vs[i] = v[i]; xs[i] = (IdxT)i; }
        v = vs;
        x = xs;
    } else {
        x = (IdxT *)(A.idx + base);                          // only reached with IdxT = int32_t
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) x[i] = (IdxT)i;
    }
    int32_t *Lbase = A.scr + 2 * base, *Rbase = Lbase + len;        // a range [pl, pr] uses the slice [pl, pr] of both lists
    if (threadIdx.x == 0) {
        st_lo[0] = 0; st_hi[0] = len - 1; st_d[0] = 2 * (31 - __clz((unsigned)len));
        s_sp = 1; s_pending = 1; s_lock = 0;
    }
    __syncthreads();
    volatile int32_t *vsp = &s_sp, *vpend = &s_pending;
    auto lock = [&]() { while (atomicCAS(&s_lock, 0, 1) != 0) __nanosleep(20); __threadfence_block(); };
    auto unlock = [&]() { __threadfence_block(); atomicExch(&s_lock, 0); };
    int nsmall = 0;
    // sub-ranges of at most ICL_NPS_LANE elements are collected per warp and sorted 32 at a time, one per lane
    auto flush_small = [&]() {
        __syncwarp();
        if (lane < nsmall) nps_seq(v, x, sm_lo[w][lane], sm_hi[w][lane], sm_d[w][lane], sm_p[w][lane] != 0);
        nsmall = 0;
        __syncwarp();
    };
    for (;;) {
        // ---- pop a range (lane 0), or leave when the stack is empty and nobody is still producing
        int32_t pl = 0, pr = 0, cdepth = 0, got = 0;
        if (lane == 0) {
            for (;;) {
                lock();
                if (*vsp > 0) {
                    const int32_t t = *vsp - 1;
                    pl = st_lo[t]; pr = st_hi[t]; cdepth = st_d[t];
                    *vsp = t;
                    got = 1;
                    unlock();
                    break;
                }
                const int32_t pend = *vpend;
                unlock();
                if (pend == 0) break;
                __nanosleep(200);
            }
        }
        got = __shfl_sync(FULL, got, 0);
        if (!got) break;
        pl = __shfl_sync(FULL, pl, 0); pr = __shfl_sync(FULL, pr, 0); cdepth = __shfl_sync(FULL, cdepth, 0);
        bool popped = true;
        if (cdepth < 0 && pr - pl + 1 > ICL_NPS_LANE) {
            if (lane == 0) nps_heapsort(v, x + pl, pr - pl + 1);
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicSub(&s_pending, 1); }
            continue;
        }
        while (pr - pl + 1 > ICL_NPS_LANE) {
            int32_t *Ls = Lbase + pl, *Rs = Rbase + pl;
            // median of three with index swaps, pivot parked at pr - 1 (lane 0; a handful of dependent loads)
            int32_t vp = 0;
            if (lane == 0) {
                const int32_t pm = pl + ((pr - pl) >> 1);
                int32_t a = x[pl], b = x[pm], c = x[pr];
                if (v[b] < v[a]) { int32_t t = a; a = b; b = t; }
                if (v[c] < v[b]) { int32_t t = b; b = c; c = t; }
                if (v[b] < v[a]) { int32_t t = a; a = b; b = t; }
                vp = v[b];
                x[pl] = (IdxT)a; x[pr] = (IdxT)c;
                x[pm] = x[pr - 1];                  // INTP_SWAP(*pm, *pj) with pj = pr - 1 (pm < pr - 1 here)
                x[pr - 1] = (IdxT)b;
            }
            vp = __shfl_sync(FULL, vp, 0);
            __syncwarp();
            // stop lists: L = positions in [pl + 1, pr - 1] with key >= pivot, ascending;
            //             R = positions in [pl, pr - 2] with key <= pivot — listed ascending in the same pass and read
            //             back to front below
            int32_t nL = 0, nR = 0;
            for (int32_t b0 = pl; b0 <= pr - 1; b0 += 32) {
                const int32_t q = b0 + lane;
                const int32_t kq = q <= pr - 1 ? v[x[q]] : 0;
                const bool fl = q >= pl + 1 && q <= pr - 1 && !(kq < vp);
                const bool fr = q <= pr - 2 && !(vp < kq);
                const uint32_t ml = __ballot_sync(FULL, fl), mr = __ballot_sync(FULL, fr);
                const uint32_t below = (1u << lane) - 1u;
                if (fl) Ls[nL + __popc(ml & below)] = q;
                if (fr) Rs[nR + __popc(mr & below)] = q;
                nL += __popc(ml);
                nR += __popc(mr);
            }
            __syncwarp();
            // K = number of k with L[k] < R[k] (a prefix: L ascends, R descends); then swap those pairs
            const int32_t kmax = nL < nR ? nL : nR;
            int32_t K = 0;
            for (int32_t b0 = 0; b0 < kmax; b0 += 32) {
                const int32_t k = b0 + lane;
                const bool f = k < kmax && Ls[k] < Rs[nR - 1 - k];
                const uint32_t m = __ballot_sync(FULL, f);
                K += __popc(m);
                if (m != FULL) break;
            }
            for (int32_t k = lane; k < K; k += 32) {
                const int32_t a = Ls[k], b = Rs[nR - 1 - k];
                const IdxT t = x[a]; x[a] = x[b]; x[b] = t;
            }
            __syncwarp();
            // where the left scan stops for good: the next stop of the untouched array, or the slot the last swap filled
            // from the left side, whichever comes first
            int32_t pi = Ls[K];
            if (K >= 1 && Rs[nR - K] < pi) pi = Rs[nR - K];
            if (lane == 0) { const IdxT t = x[pi]; x[pi] = x[pr - 1]; x[pr - 1] = t; }
            __syncwarp();
            --cdepth;
            // push the larger side for any warp to take, continue with the smaller one (both at the decremented depth)
            int32_t ql, qr;
            if (pi - pl < pr - pi) { ql = pi + 1; qr = pr; pr = pi - 1; }
            else { ql = pl; qr = pi - 1; pl = pi + 1; }
            if (lane == 0) {
                lock();
                const int32_t t = *vsp;
                st_lo[t] = ql; st_hi[t] = qr; st_d[t] = cdepth;
                *vsp = t + 1;
                atomicAdd(&s_pending, 1);             // (the decrements are atomics outside the lock)
                unlock();
            }
            popped = false;
            __syncwarp();
        }
        if (pr > pl) {                                       // the rest of this sub-range goes to one lane, with its depth
            if (lane == 0) { sm_lo[w][nsmall] = pl; sm_hi[w][nsmall] = pr; sm_d[w][nsmall] = cdepth; sm_p[w][nsmall] = popped ? 1 : 0; }
            nsmall++;
            if (nsmall == 32) flush_small();
        }
        if (lane == 0) { __threadfence_block(); atomicSub(&s_pending, 1); }
    }
    if (nsmall) flush_small();
    __syncthreads();
    if (fits) {
        int32_t *out = A.idx + base;
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) out[i] = x[i];
    }
}
#endif

// ------------------------------------------------------------------------------------------------
// numpy's AVX-512 and AVX2 argsort, reproduced exactly on the device.
//
// On CPUs with AVX-512 numpy.argsort(kind='quicksort') on int64 is x86-simd-sort's `avx512_argsort` (this pool's hosts),
// on other x86 CPUs its AVX2 counterpart.  The algorithm (restated from the library's source as I know it and checked
// step by step against numpy 2.3.5 — tools/xss_emulation.py: 8-lane network, 16, 32..256 elements, pivot + partition,
// std::sort fallback, sorted / constant arrays, real DP score arrays — before any CUDA was written; described for 8
// lanes, the AVX2 differences are at the kernel):
//   keys already in non-decreasing order: the identity (top level only).
//   argsort_(left, right, iters = 2 * floor(log2 n)):
//     iters <= 0            -> std::sort of the index range with the comparator arr[a] < arr[b] (libstdc++ introsort)
//     at most 256 elements  -> key-index bitonic network over 8-lane registers (padding key = max, padding index = 0):
//                              every stage pairs a position with a partner, keeps min or max, and the index follows the
//                              rule  new_index = own if new_key == own_key else partner's
//     else pivot = 5th smallest of the 8 keys at left + k * ((right - left) / 8), k = 1..8; partition (below);
//          recurse left if pivot != smallest key seen, right if pivot != largest key seen, both with iters - 1.
//   partition (4 x 8 lanes unrolled; `right` exclusive): the first (right - left) % 32 elements one by one
//   (key >= pivot: swap with --right, else ++left); the first and the last 32 elements are set aside; then groups of
//   32 are taken from the right if fewer elements have been stored on the right than on the left, else from the left;
//   of every 8-lane vector the keys < pivot are compress-stored at l_store, the keys >= pivot below r_store + 8; the two
//   groups set aside go last (left one first).
// Every step is a function of the keys alone, so a warp can do the 32 lanes of a group, or the positions of a network
// stage, at once, and disjoint sub-ranges can be taken by different warps in any order, without changing the result.
// DP score arrays always reach the std::sort fallback: the mass of never-updated points holds two or three distinct
// keys, a pivot equal to the smallest key moves everything and makes no progress, and the budget runs out.  Those
// passes are applied in closed form (XsPerm below); std::sort itself (median of three to the front, unguarded Hoare
// partition, 16-element threshold, final insertion sort) is done by a warp with stop lists down to ICL_XS_LANE elements
// and by single lanes below that; its heap-sort branch (depth limit) has never been observed and hands the segment
// back to the host (flag), as does anything else this kernel does not cover.
// ------------------------------------------------------------------------------------------------
#ifndef ICL_EMU
struct XsArgs {
    const int32_t *key;        // [2n] segments back to back: the REVERSED score arrays
    int32_t *idx;              // [2n] out: argsort of every segment (segment-local indices)
    const int64_t *off;        // [segments + 1] start of every segment in key / idx
    const int32_t *order;      // segment order (longest first)
    int32_t *host_flag;        // [segments] set to 1 when a segment must be re-sorted by numpy on the host
    int32_t *scr;              // [4n] scratch: stop lists of the std::sort partitions
    int32_t smem_cap;
    int32_t W;                 // lanes of the emulated vector: 8 = AVX-512, 4 = AVX2
    int32_t first;             // this launch covers order[first .. first + gridDim.x)
    int32_t keys_in_smem;      // keys next to the indices in shared memory (else only the indices; keys through L1)
    int32_t lane_max;          // std::sort pieces of at most this many elements are sorted by one lane each
    int32_t force_giveup;      // test hook (env ICL_B200_XS_FORCE_GIVEUP): the first segment of the launch gives up at once
    int32_t coop_min;          // ranges longer than this are processed by the whole CTA (phase 1)
    uint32_t *chunk;           // [2 (total / 32 + 3 segments)] per-chunk stop counts of very long std::sort ranges
};

// The key-index network is a bitonic sort with flip merges over P = 8 * numVecs slots: for S = 2, 4, .., P: compare
// slot i with S - 1 - i inside every group of S (min to the lower slot), then for m = S/4, S/8, .., 1 compare i with
// i + m (bit m of i clear).  One thread per pair; the index follows its key, and on equal keys nothing moves.
__device__ __forceinline__ void xs_pair(int2 *nkv, int i, int j) {      // .x = key, .y = index
    const int2 a = nkv[i], b = nkv[j];
    if (b.x < a.x) {                                         // strictly out of order: swap key and index
        nkv[i] = b;
        nkv[j] = a;
    }
}

template <class IdxT>
__device__ void xs_network(const int32_t *v, IdxT *x, int32_t left, int32_t n, int2 *nkv, int lane) {
    int P = 256;
    while (P > 8 && n * 2 <= P) P >>= 1;
    for (int i = lane; i < P; i += 32) {
        if (i < n) { const int32_t a = (int32_t)x[left + i]; nkv[i] = make_int2(v[a], a); }
        else nkv[i] = make_int2(0x7fffffff, 0);
    }
    __syncwarp();
    const int half = P >> 1;
    for (int S = 2; S <= P; S <<= 1) {
        for (int t = lane; t < half; t += 32) {              // flip stage: (G*S + k, G*S + S-1-k), k < S/2
            const int hs = S >> 1, G = t / hs, k = t - G * hs;
            xs_pair(nkv, G * S + k, G * S + S - 1 - k);
        }
        __syncwarp();
        for (int m = S >> 2; m >= 1; m >>= 1) {
            for (int t = lane; t < half; t += 32) {          // xor stage: insert a zero bit at log2(m)
                const int i = ((t & ~(m - 1)) << 1) | (t & (m - 1));
                xs_pair(nkv, i, i | m);
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < n; i += 32) x[left + i] = (IdxT)nkv[i].y;
    __syncwarp();
}

// ---- libstdc++ std::sort(x + first, x + last, [v](a, b) { return v[a] < v[b]; }), the fallback of argsort_ ----
// __introsort_loop on [f, l) with the given depth budget by ONE lane (sub-ranges of at most ICL_XS_LANE elements),
// followed by the part of __final_insertion_sort that concerns this sub-range: the loop leaves pieces of at most 16
// elements that are ordered among each other, and the final insertion sort never moves an element across a piece
// boundary (it stops at the first element that is not greater), so it is an insertion sort of every piece — here of
// the whole sub-range, which is the same.  Returns false at the depth limit (heap sort: not reproduced).
#define ICL_XS_LANE 96
template <class IdxT>
__device__ bool xs_std_seq(const int32_t *v, IdxT *x, int32_t f, int32_t l, int depth) {
    const int32_t first = f, last = l;
    int32_t sf[12], sl[12];
    int sd[12], sp = 0;
    for (;;) {
        while (l - f > 16) {
            if (depth == 0) return false;
            --depth;
            const int32_t mid = f + ((l - f) >> 1), a = f + 1, b = mid, c = l - 1;
            const int32_t ka = v[x[a]], kb = v[x[b]], kc = v[x[c]];
            int32_t r;
            if (ka < kb) { if (kb < kc) r = b; else if (ka < kc) r = c; else r = a; }
            else if (ka < kc) r = a;
            else if (kb < kc) r = c;
            else r = b;
            IdxT t = x[f]; x[f] = x[r]; x[r] = t;
            const int32_t pv = v[x[f]];
            int32_t i = f + 1, j = l;
            for (;;) {
                while (v[x[i]] < pv) ++i;
                --j;
                while (pv < v[x[j]]) --j;
                if (!(i < j)) break;
                t = x[i]; x[i] = x[j]; x[j] = t;
                ++i;
            }
            // left part later, right part now (order is free); a part of at most 16 elements is finished (the
            // insertion sort below covers it), so at most 96 / 17 parts wait here at any time
            if (i - f > 16) { sf[sp] = f; sl[sp] = i; sd[sp] = depth; sp++; }
            f = i;
        }
        if (sp == 0) break;
        sp--;
        f = sf[sp]; l = sl[sp]; depth = sd[sp];
    }
    for (int32_t i = first + 1; i < last; ++i) {              // insertion sort (stable) of the sub-range
        const IdxT val = x[i];
        const int32_t kv = v[val];
        int32_t q = i;
        while (q > first && kv < v[x[q - 1]]) { x[q] = x[q - 1]; --q; }
        x[q] = val;
    }
    return true;
}

// std::sort on a range of EQUAL keys is a fixed permutation of positions: no comparison is ever true, so every level of
// __introsort_loop moves the middle element to the front (__move_median_to_first picks `mid`), __unguarded_partition
// swaps (f + 1 + k, l - 1 - k) for k < (l - f - 1) / 2 — the reversal of [f + 1, l) — and cuts at f + 1 + (l - f - 1) / 2;
// the final insertion sort moves nothing.  Where the element at q ends up (tools/xss_emulation.py `std_sort_equal_pos`,
// checked against the emulation of std::sort for all sizes 1..400 and some larger ones):
__device__ __forceinline__ int32_t xs_std_equal_pos(int32_t f, int32_t l, int32_t q) {
    while (l - f > 16) {
        const int32_t mid = f + ((l - f) >> 1);
        q = (q == f) ? f + l - mid : (q == mid) ? f : f + l - q;
        const int32_t cut = f + 1 + ((l - f - 1) >> 1);
        if (q < cut) l = cut; else f = cut;
    }
    return q;
}
// levels of __introsort_loop that a range of s equal keys goes through (its larger part decides)
__device__ __forceinline__ int xs_std_equal_levels(int32_t s) {
    int lv = 0;
    while (s > 16) {
        const int32_t a = 1 + ((s - 1) >> 1), b = s - a;
        s = a > b ? a : b;
        lv++;
    }
    return lv;
}

#define ICL_XS_WARPS_DEFAULT 8
#define ICL_XS_SMEM_MAX 208768     // dynamic shared memory a CTA may ask for: 227 KB minus the static arrays of 8 warps
#define ICL_XS_STACK 256
#define ICL_XS_BIG 64              // ranges waiting for the cooperative phase (depth of the recursion, at most)

// A partition pass whose pivot equals the smallest key of its range moves every element to the "right" side: nothing
// is decided by the keys, the pass is a fixed permutation of the positions [left, R) and the recursion comes back to the
// same range with one iteration less.  DP score arrays spend most of their budget this way (the never-updated points
// are a mass of equal smallest keys), so the permutation is written in closed form and a run of t such passes is applied
// at once.  One pass: the first pre = (R - left) mod G elements rotate through position `left` to the end of the
// range; of the M groups of G that remain, group 1 is stored first, then M-2, M-3, .., 2, then the two groups set aside
// (0 and M-1); the k-th stored group fills [rend - (k+1)G, rend - kG) from the top, vector by vector (AVX2: lanes
// reversed).  Checked against the sequential partition in tools/xss_emulation.py (`check_fast_forward`).
// Groups 2 .. M-2 keep their place and only reverse their vectors (and lanes, AVX2): on them the pass is an involution;
// groups 0, 1, M-1 and the rotated elements move among themselves.
struct XsPerm {
    int32_t left, R, rend, M, pre;
};
template <int W>
__device__ __forceinline__ int32_t xs_fwd(const XsPerm &P, int32_t q) {          // where the element at q goes
    if (P.pre) {
        if (q == P.left) q = P.R - 1;
        else if (q >= P.rend) q = (q == P.rend) ? P.left : q - 1;
    }
    if (q >= P.rend) return q;
    const int32_t e = q - P.left;
    constexpr int wsh = W == 8 ? 3 : 2;
    const int32_t g = e >> (wsh + 2), r = e & (4 * W - 1);
    const int32_t j = r >> wsh, t = r & (W - 1);
    int32_t k;
    if (P.M == 2) k = g;
    else if (g == 1) k = 0;
    else if (g == 0) k = P.M - 2;
    else if (g == P.M - 1) k = P.M - 1;
    else k = P.M - 1 - g;
    const int32_t sl = W == 8 ? t : W - 1 - t;
    return P.rend - k * 4 * W - (j + 1) * W + sl;
}
template <int W>
__device__ __forceinline__ int32_t xs_fwd_t(const XsPerm &P, int32_t q, int t) {  // after t passes
    constexpr int wsh = W == 8 ? 3 : 2;
    const int32_t e = q - P.left;
    const int32_t g = e >> (wsh + 2);
    if (q < P.rend && g >= 2 && g <= P.M - 2) {
        if (!(t & 1)) return q;
        const int32_t r = e & (4 * W - 1);
        const int32_t j = r >> wsh, l = r & (W - 1);
        return P.left + g * 4 * W + (3 - j) * W + (W == 8 ? l : W - 1 - l);
    }
    for (int k = 0; k < t; k++) q = xs_fwd<W>(P, q);
    return q;
}
template <int W>
__device__ __forceinline__ int32_t xs_inv(const XsPerm &P, int32_t d) {          // where the element at d came from
    constexpr int wsh = W == 8 ? 3 : 2;
    if (d >= P.rend) return d == P.R - 1 ? P.left : d + 1;
    const int32_t u = P.rend - 1 - d;
    const int32_t k = u >> (wsh + 2), r = u & (4 * W - 1);
    const int32_t j = r >> wsh, sl = W - 1 - (r & (W - 1));
    const int32_t t = W == 8 ? sl : W - 1 - sl;
    int32_t g;
    if (P.M == 2) g = k;
    else if (k == 0) g = 1;
    else if (k == P.M - 2) g = 0;
    else if (k == P.M - 1) g = P.M - 1;
    else g = P.M - 1 - k;
    int32_t q = P.left + g * 4 * W + j * W + t;
    if (P.pre && q == P.left) q = P.rend;
    return q;
}
// One CTA of NW warps per segment; the warps take ranges from a shared stack (sub-ranges are disjoint and carry their
// own budgets, so the order is free), each with its own network scratch and its own list of small std::sort
// sub-ranges.  A.W = lanes of the emulated vector: 8 (AVX-512) or 4 (AVX2: same algorithm on 4-lane vectors, groups of 16,
// pivot = 3rd smallest of 4 samples, and — its compress-store is a permute + full-width store — the keys >= pivot of a
// vector land in REVERSED lane order).
template <class IdxT, int NW, int W>
__global__ void __launch_bounds__(32 * NW) k_xss_argsort(XsArgs A) {
    __shared__ int2 nkv_all[NW][256];
    __shared__ int32_t st_lo[ICL_XS_STACK], st_hi[ICL_XS_STACK], st_it[ICL_XS_STACK];   // st_it >= 0: argsort_ range with that
                                                              // budget; < 0: std::sort range [lo, hi), depth budget -st_it - 1
    __shared__ int32_t sm_lo_all[NW][32], sm_hi_all[NW][32], sm_d_all[NW][32];
    __shared__ int32_t s_sp, s_pending, s_lock, s_giveup;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int2 *nkv = nkv_all[wid];
    int32_t *sm_lo = sm_lo_all[wid], *sm_hi = sm_hi_all[wid], *sm_d = sm_d_all[wid];
    const uint32_t FULL = 0xffffffffu;
    constexpr int G = 4 * W;                                  // W lanes per vector, G elements per unrolled group
    constexpr int wsh = W == 8 ? 3 : 2;
    const int32_t it = A.order[A.first + blockIdx.x];
    const int64_t base = A.off[it];
    const int32_t len = (int32_t)(A.off[it + 1] - base);
    if (len <= 0) return;
    extern __shared__ int32_t xs_sm[];
    const bool fits = sizeof(IdxT) == 2 || len <= A.smem_cap;   // indices in shared memory
    const bool kin = fits && A.keys_in_smem;                     // keys too
    const int32_t *v = A.key + base;
    IdxT *x;
    IdxT *xsh = (IdxT *)(xs_sm + (A.keys_in_smem ? A.smem_cap : 0));
    if (fits) {
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) xsh[i] = (IdxT)i;
        if (kin) {
            // keys to shared memory with 16-byte global loads (one CTA per SM in the large classes: nothing else hides them)
            const int32_t head = (int32_t)((4 - (base & 3)) & 3) < len ? (int32_t)((4 - (base & 3)) & 3) : len;
            const int32_t nvec = (len - head) >> 2;
            const int4 *src = (const int4 *)(v + head);
            if ((int32_t)threadIdx.x < head) xs_sm[threadIdx.x] = v[threadIdx.x];
            for (int32_t j = threadIdx.x; j < nvec; j += blockDim.x) {
                const int4 q = __ldg(src + j);
                int32_t *d = xs_sm + head + 4 * j;
                d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            }
            for (int32_t i = head + 4 * nvec + threadIdx.x; i < len; i += blockDim.x) xs_sm[i] = v[i];
        }
    } else {
        x = (IdxT *)(A.idx + base);
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) x[i] = (IdxT)i;
    }
    __syncthreads();
    // keys already in non-decreasing order: the library leaves the identity (top level only)
    int uns = 0;
    {
        const int32_t *vv = kin ? xs_sm : v;
        for (int32_t i = threadIdx.x; i + 1 < len; i += blockDim.x) uns |= vv[i] > vv[i + 1];
    }
    const int unsorted = __syncthreads_or(uns);
    if (fits) {
        if (kin) v = xs_sm;
        x = xsh;
    }
    int32_t *Lbase = A.scr + 2 * base, *Rbase = Lbase + len;  // stop lists of the std::sort partitions (slice [f, l) per range)
    if (threadIdx.x == 0) {
        st_lo[0] = 0; st_hi[0] = len - 1; st_it[0] = 2 * (31 - __clz((unsigned)len));
        s_sp = unsorted ? 1 : 0; s_pending = unsorted ? 1 : 0; s_lock = 0;
        s_giveup = (A.force_giveup && blockIdx.x == 0) ? 1 : 0;
    }
    __syncthreads();
    // ------------------------------------------------------------------------------------------------------------
    // phase 1: ranges longer than A.coop_min, the whole CTA on one range at a time.
    // A long segment used to spend its top levels on ONE warp (a partition step per 32 keys, each a dependent chain);
    // everything that warp does is a function of per-group counts only, so it can be laid out in advance:
    //  * partition: with lt(g) = keys < pivot in group g, the side rule "take the next group from the right iff fewer
    //    elements have been stored on the right than on the left" reads  sum of lt over the groups taken from the right <
    //    sum of (G - lt) over the groups taken from the left.  Both sums only grow, so the order in which groups are taken is
    //    the MERGE of the two prefix-sum sequences A (from the left) and B (from the right), A first on ties: a group's turn
    //    and the two store pointers at its turn follow from A, B and one binary search per group
    //    (checked against the sequential partition: tools/xss_emulation.py `partition_merge_path`).  All warps count, two
    //    scan, all scatter out of place;
    //  * passes without progress: a fixed permutation of positions, applied by all threads;
    //  * std::sort's unguarded partition: stop lists by all warps (counts per 32 positions, scan, fill), crossing point by
    //    counting, swaps in parallel.
    // Sub-ranges of at most coop_min elements go to the shared stack for phase 2 (one warp per range, as before).
    // ------------------------------------------------------------------------------------------------------------
    __shared__ int32_t bg_lo[ICL_XS_BIG], bg_hi[ICL_XS_BIG], bg_it[ICL_XS_BIG];
    __shared__ int32_t s_nbig, c_a, c_b, c_c, c_d, c_min, c_max, c_cnt0, c_cnt1;
    {
        const int32_t coop_min = A.coop_min > 512 ? A.coop_min : 512;
        const int nthr = 32 * NW;
        if (threadIdx.x == 0) {
            s_nbig = 0;
            if (unsorted && len > coop_min) {                  // the whole segment starts here
                bg_lo[0] = st_lo[0]; bg_hi[0] = st_hi[0]; bg_it[0] = st_it[0];
                s_nbig = 1; s_sp = 0; s_pending = 0;
            }
        }
        __syncthreads();
        // thread 0 only, between barriers
        auto place = [&](int32_t lo_, int32_t hi_, int32_t it_) {
            const int32_t cnt = it_ < 0 ? hi_ - lo_ : hi_ + 1 - lo_;
            if (cnt > coop_min && s_nbig < ICL_XS_BIG) { bg_lo[s_nbig] = lo_; bg_hi[s_nbig] = hi_; bg_it[s_nbig] = it_; s_nbig++; }
            else if (s_sp < ICL_XS_STACK) { st_lo[s_sp] = lo_; st_hi[s_sp] = hi_; st_it[s_sp] = it_; s_sp++; s_pending++; }
            else s_giveup = 1;
        };
        uint32_t *const smA = (uint32_t *)&nkv_all[0][0];     // prefix sums of the partition: free while no network runs
        constexpr int smcap = NW * 256;                        // uint32 entries per array in nkv_all (two arrays)
        for (;;) {
            __syncthreads();
            if (s_nbig == 0 || s_giveup) break;
            int32_t left = bg_lo[s_nbig - 1], right = bg_hi[s_nbig - 1], iters = bg_it[s_nbig - 1];
            __syncthreads();
            if (threadIdx.x == 0) s_nbig--;
            if (iters < 0) {
                // ---------------------------------------------------------------- std::sort range [f, l), one partition step
                const int32_t f = left, l = right;
                int depth = -iters - 1;
                if (depth == 0) { if (threadIdx.x == 0) s_giveup = 1; continue; }     // heap-sort branch: the host ranks
                --depth;
                if (threadIdx.x == 0) {                      // __move_median_to_first(f, f + 1, mid, l - 1)
                    const int32_t mid = f + ((l - f) >> 1), a = f + 1, b = mid, c = l - 1;
                    const int32_t ka = v[x[a]], kb = v[x[b]], kc = v[x[c]];
                    int32_t r;
                    if (ka < kb) { if (kb < kc) r = b; else if (ka < kc) r = c; else r = a; }
                    else if (ka < kc) r = a;
                    else if (kb < kc) r = c;
                    else r = b;
                    const IdxT t = x[f]; x[f] = x[r]; x[r] = t;
                    c_a = v[x[f]];
                    c_cnt0 = 0; c_cnt1 = 0;
                }
                __syncthreads();
                const int32_t pv = c_a;
                // stop lists of [f + 1, l): per chunk of 32 positions the counts, a scan over chunks, then the lists
                int32_t *Ls = Lbase + f, *Rs = Rbase + f;
                const int32_t nch = (l - (f + 1) + 31) >> 5;
                uint32_t *chL, *chR;                          // exclusive chunk offsets
                if (nch + 1 <= smcap) { chL = smA; chR = smA + smcap; }
                else { chL = A.chunk + 2 * ((base >> 5) + 3 * (int64_t)it); chR = chL + nch + 1; }   // the segment's own region
                {
                    int32_t kmin = pv, kmax = pv;
                    for (int32_t ch = wid; ch < nch; ch += NW) {
                        const int32_t q = f + 1 + (ch << 5) + lane;
                        const int32_t kq = q < l ? v[x[q]] : pv;
                        const uint32_t ml = __ballot_sync(FULL, q < l && !(kq < pv)), mr = __ballot_sync(FULL, q < l && !(pv < kq));
                        kmin = kq < kmin ? kq : kmin;
                        kmax = kq > kmax ? kq : kmax;
                        if (lane == 0) { chL[ch] = __popc(ml); chR[ch] = __popc(mr); }
                    }
                    const bool eq = kmin == kmax;
                    if (!__all_sync(FULL, eq)) c_cnt1 = 1;    // (benign race: every writer stores 1)
                }
                __syncthreads();
                if (c_cnt1 == 0 && depth + 1 >= xs_std_equal_levels(l - f)) {
                    // every key of the range equals the pivot: std::sort is a fixed permutation here (xs_std_equal_pos), all
                    // levels at once.  The median swap of this level is undone first: the permutation starts from the range
                    // as std::sort found it.
                    if (threadIdx.x == 0) { const int32_t mid = f + ((l - f) >> 1); const IdxT t = x[f]; x[f] = x[mid]; x[mid] = t; }
                    __syncthreads();
                    int32_t *tmp = Lbase + f;
                    for (int32_t q = f + threadIdx.x; q < l; q += nthr) tmp[xs_std_equal_pos(f, l, q) - f] = (int32_t)x[q];
                    __syncthreads();
                    for (int32_t q = f + threadIdx.x; q < l; q += nthr) x[q] = (IdxT)tmp[q - f];
                    continue;
                }
                if (wid == 0 || (NW > 1 && wid == 1)) {       // exclusive scans over the chunks (warp 0: left, warp 1: right)
                    for (int pass = 0; pass < (NW > 1 ? 1 : 2); pass++) {
                        uint32_t *arr = (NW > 1 ? wid : pass) == 0 ? chL : chR;
                        uint32_t carry = 0;
                        for (int32_t b0 = 0; b0 <= nch; b0 += 32) {
                            const int32_t i = b0 + lane;
                            uint32_t val = i < nch ? arr[i] : 0u, inc = val;
                            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
                            if (i <= nch) arr[i] = carry + inc - val;
                            carry += __shfl_sync(FULL, inc, 31);
                        }
                    }
                }
                __syncthreads();
                const int32_t nL = (int32_t)chL[nch], nR = (int32_t)chR[nch];
                for (int32_t ch = wid; ch < nch; ch += NW) {
                    const int32_t q = f + 1 + (ch << 5) + lane;
                    const int32_t kq = q < l ? v[x[q]] : 0;
                    const bool fl = q < l && !(kq < pv), fr = q < l && !(pv < kq);
                    const uint32_t ml = __ballot_sync(FULL, fl), mr = __ballot_sync(FULL, fr);
                    const uint32_t below = (1u << lane) - 1u;
                    if (fl) Ls[chL[ch] + __popc(ml & below)] = q;
                    if (fr) Rs[chR[ch] + __popc(mr & below)] = q;
                }
                __syncthreads();
                // K = number of k with L[k] < R[nR - 1 - k] (a prefix); swap those pairs
                const int32_t kmax2 = nL < nR ? nL : nR;
                int32_t myK = 0;
                for (int32_t k = threadIdx.x; k < kmax2; k += nthr) myK += Ls[k] < Rs[nR - 1 - k];
                myK = __reduce_add_sync(FULL, myK);
                if (lane == 0 && myK) atomicAdd(&c_cnt0, myK);
                __syncthreads();
                const int32_t K = c_cnt0;
                for (int32_t k = threadIdx.x; k < K; k += nthr) {
                    const int32_t a = Ls[k], b = Rs[nR - 1 - k];
                    const IdxT t = x[a]; x[a] = x[b]; x[b] = t;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int32_t cut = K < nL ? Ls[K] : l;
                    if (K >= 1 && Rs[nR - K] < cut) cut = Rs[nR - K];
                    // right part [cut, l) and left part [f, cut), both with the decremented depth; parts of at most 16
                    // elements are left to the final insertion sort of the piece they end up in: they ride with a neighbour
                    if (l - cut > 1) place(cut, l, -depth - 1);
                    if (cut - f > 1) place(f, cut, -depth - 1);
                }
                continue;
            }
            if (iters == 0) {                                 // argsort_ out of budget: std::sort of the range
                if (threadIdx.x == 0) {
                    const int32_t cnt = right + 1 - left;
                    place(left, right + 1, -(2 * (31 - __clz((unsigned)cnt))) - 1);
                }
                continue;
            }
            // ---------------------------------------------------------------- argsort_ range [left, right], > 256 elements
            const int32_t ssize = (right - left) >> wsh;
            if (wid == 0) {
                const int32_t mine0 = lane < W ? v[x[left + (lane + 1) * ssize]] : 0x7fffffff;
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int32_t o = __shfl_sync(FULL, mine0, j);
                    rank += j < W && ((o < mine0) || (o == mine0 && j < lane));
                }
                const uint32_t who = __ballot_sync(FULL, lane < W && rank == (W >> 1));
                const int32_t pv = __shfl_sync(FULL, mine0, __ffs(who) - 1);
                const int32_t mn = __reduce_min_sync(FULL, mine0);
                if (lane == 0) { c_a = pv; c_b = mn; c_min = 0x7fffffff; c_max = (int32_t)0x80000000; c_cnt0 = 0; }
            }
            __syncthreads();
            int32_t pivot = c_a;
            if (c_b == pivot) {
                // the pivot is the smallest sample: if it is the smallest key of the range, the pass decides nothing
                int32_t rmin = 0x7fffffff, rmax = (int32_t)0x80000000;
                for (int32_t i = left + threadIdx.x; i <= right; i += nthr) {
                    const int32_t kv = v[x[i]];
                    rmin = kv < rmin ? kv : rmin;
                    rmax = kv > rmax ? kv : rmax;
                }
                rmin = __reduce_min_sync(FULL, rmin);
                rmax = __reduce_max_sync(FULL, rmax);
                if (lane == 0) { atomicMin(&c_min, rmin); atomicMax(&c_max, rmax); }
                __syncthreads();
                rmin = c_min; rmax = c_max;
                if (pivot == rmin) {
                    XsPerm P;
                    P.left = left; P.R = right + 1;
                    P.pre = (P.R - left) & (G - 1);
                    P.rend = P.R - P.pre;
                    P.M = (P.rend - left) >> (wsh + 2);
                    if (wid == 0) {                            // how many such passes follow each other, and what ends them
                        int32_t q = lane < W ? left + (lane + 1) * ssize : left;
                        int t = 0, state;
                        int32_t pv = pivot;
                        for (;;) {
                            ++t;
                            if (pv == rmax) { state = 0; break; }
                            if (iters - t == 0) { state = 1; break; }
                            q = xs_inv<W>(P, q);
                            const int32_t mine = lane < W ? v[x[q]] : 0x7fffffff;
                            int rank = 0;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int32_t o = __shfl_sync(FULL, mine, j);
                                rank += j < W && ((o < mine) || (o == mine && j < lane));
                            }
                            const uint32_t who = __ballot_sync(FULL, lane < W && rank == (W >> 1));
                            pv = __shfl_sync(FULL, mine, __ffs(who) - 1);
                            if (pv != rmin) { state = 2; break; }
                        }
                        if (lane == 0) { c_a = pv; c_c = t; c_d = state; }
                    }
                    __syncthreads();
                    const int t = c_c, state = c_d;
                    pivot = c_a;
                    int32_t *tmp = Lbase + left;               // out of place through the scratch slice
                    for (int32_t i = left + threadIdx.x; i < P.R; i += nthr) tmp[xs_fwd_t<W>(P, i, t) - left] = (int32_t)x[i];
                    __syncthreads();
                    for (int32_t i = left + threadIdx.x; i < P.R; i += nthr) x[i] = (IdxT)tmp[i - left];
                    __syncthreads();
                    iters -= t;
                    if (state == 0) continue;
                    if (state == 1) {
                        if (threadIdx.x == 0) {
                            const int32_t cnt = right + 1 - left;
                            place(left, right + 1, -(2 * (31 - __clz((unsigned)cnt))) - 1);
                        }
                        continue;
                    }
                    if (threadIdx.x == 0) { c_min = 0x7fffffff; c_max = (int32_t)0x80000000; }
                    __syncthreads();
                }
            }
            // ---- partition [left, right + 1) around `pivot`
            const int32_t range_left = left, range_right = right;
            if (threadIdx.x == 0) {
                int32_t lf = left, rend = right + 1, kmin = 0x7fffffff, kmax = (int32_t)0x80000000;
                for (int32_t pre = (rend - lf) & (G - 1); pre > 0; --pre) {
                    const int32_t kv = v[x[lf]];
                    kmin = kv < kmin ? kv : kmin;
                    kmax = kv > kmax ? kv : kmax;
                    if (!(kv < pivot)) { --rend; const IdxT t = x[lf]; x[lf] = x[rend]; x[rend] = t; }
                    else ++lf;
                }
                c_a = lf; c_b = rend; c_min = kmin; c_max = kmax;
            }
            __syncthreads();
            const int32_t L0 = c_a, R0 = c_b;
            const int32_t M = (R0 - L0) >> (wsh + 2), m = M - 2;          // groups; the first and the last are set aside
            uint32_t *PA, *PB;                                // PA[i] = keys >= pivot in the first i middle groups from the left,
            if (M <= smcap) { PA = smA; PB = smA + smcap; }   // PB[j] = keys < pivot in the first j from the right (i, j = 0..m)
            else { PA = (uint32_t *)(Rbase + left); PB = PA + M; }
            const bool act = lane < G;                        // AVX2: 16 lanes carry a group
            {
                int32_t kmin = 0x7fffffff, kmax = (int32_t)0x80000000;
                for (int32_t gb = wid << 5; gb < M; gb += NW << 5) {
                    int32_t myge = 0;
                    const int32_t ge_n = M - gb < 32 ? M - gb : 32;
                    for (int32_t k = 0; k < ge_n; k++) {
                        const int32_t a = act ? (int32_t)x[L0 + ((gb + k) << (wsh + 2)) + lane] : 0;
                        const int32_t kv = act ? v[a] : pivot;
                        const uint32_t mm = __ballot_sync(FULL, act && kv >= pivot);
                        kmin = kv < kmin ? kv : kmin;
                        kmax = kv > kmax ? kv : kmax;
                        if (lane == k) myge = __popc(mm);
                    }
                    const int32_t g = gb + lane;
                    if (g < M) {
                        if (g == 0) c_c = myge;
                        else if (g == M - 1) c_d = myge;
                        else { PA[g] = (uint32_t)myge; PB[M - 1 - g] = (uint32_t)(G - myge); }
                    }
                }
                kmin = __reduce_min_sync(FULL, kmin);
                kmax = __reduce_max_sync(FULL, kmax);
                if (lane == 0) { atomicMin(&c_min, kmin); atomicMax(&c_max, kmax); }
                if (threadIdx.x == 0) { PA[0] = 0; PB[0] = 0; }
            }
            __syncthreads();
            if (wid == 0 || (NW > 1 && wid == 1)) {           // inclusive scans of PA[0..m] and PB[0..m]
                for (int pass = 0; pass < (NW > 1 ? 1 : 2); pass++) {
                    uint32_t *arr = (NW > 1 ? wid : pass) == 0 ? PA : PB;
                    uint32_t carry = 0;
                    for (int32_t b0 = 0; b0 <= m; b0 += 32) {
                        const int32_t i = b0 + lane;
                        uint32_t inc = i <= m ? arr[i] : 0u;
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
                        if (i <= m) arr[i] = carry + inc;
                        carry += __shfl_sync(FULL, inc, 31);
                    }
                }
            }
            __syncthreads();
            {
                const int32_t GEm = (int32_t)PA[m], LTm = m * G - GEm;         // middle groups in total
                const int32_t ge0 = c_c;
                int32_t *tmp = Lbase + left;                  // scatter target, positions relative to L0
                const int vec = lane >> wsh, lW = lane & (W - 1);
                const uint32_t maskW = (1u << W) - 1u, belowW = (1u << lW) - 1u;
                constexpr bool rev = W == 4;
                for (int32_t gb = wid << 5; gb < M; gb += NW << 5) {
                    // the turn of group gb + lane: what has been stored on either side before it
                    const int32_t g = gb + lane;
                    int32_t LT = 0, GE = 0;
                    if (g < M) {
                        if (g == 0) { LT = LTm; GE = GEm; }
                        else if (g == M - 1) { LT = LTm + (G - ge0); GE = GEm + ge0; }
                        else {
                            int32_t i = g - 1;
                            const uint32_t Ai = PA[i];
                            int32_t lo = 0, hi = m + 1;       // j = number of j' in [0, m] with PB[j'] < Ai
                            while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (PB[mid] < Ai) lo = mid + 1; else hi = mid; }
                            int32_t j = lo;
                            if (i + j >= m) {                 // not reached from the left: taken from the right
                                j = M - 2 - g;
                                const uint32_t Bj = PB[j];
                                lo = 0; hi = m + 1;           // i = number of i' in [0, m] with PA[i'] <= Bj
                                while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (PA[mid] <= Bj) lo = mid + 1; else hi = mid; }
                                i = lo;
                            }
                            LT = i * G - (int32_t)PA[i] + (int32_t)PB[j];
                            GE = (int32_t)PA[i] + j * G - (int32_t)PB[j];
                        }
                    }
                    const int32_t ge_n = M - gb < 32 ? M - gb : 32;
                    for (int32_t k = 0; k < ge_n; k++) {
                        const int32_t l_store = __shfl_sync(FULL, LT, k), r_store = (R0 - L0) - W - __shfl_sync(FULL, GE, k);
                        const int32_t a = act ? (int32_t)x[L0 + ((gb + k) << (wsh + 2)) + lane] : 0;
                        const int32_t kv = act ? v[a] : pivot;
                        const bool isge = act && kv >= pivot;
                        const uint32_t mm = __ballot_sync(FULL, isge);
                        const int geb = __popc(mm & ((1u << (W * vec)) - 1u));   // keys >= pivot in the vectors before mine
                        const int32_t ls = l_store + W * vec - geb, rs = r_store - geb;
                        const uint32_t mv = (mm >> (W * vec)) & maskW;
                        const int ge = __popc(mv);
                        const int rk = __popc(mv & belowW);
                        const int32_t dst = isge ? (rev ? rs + W - 1 - rk : rs + W - ge + rk) : ls + __popc(~mv & belowW);
                        if (act) tmp[dst] = a;
                    }
                }
            }
            __syncthreads();
            for (int32_t i = threadIdx.x; i < R0 - L0; i += nthr) x[L0 + i] = (IdxT)(Lbase + left)[i];
            __syncthreads();
            if (threadIdx.x == 0) {
                const int32_t GEall = (int32_t)PA[m] + c_c + c_d;
                const int32_t pi = R0 - GEall;                // first key >= pivot
                const int32_t kmin = c_min, kmax = c_max;
                if (pivot != kmax) place(pi, range_right, iters - 1);
                if (pivot != kmin) place(range_left, pi - 1, iters - 1);
            }
        }
        __syncthreads();
    }
    volatile int32_t *vsp = &s_sp, *vpend = &s_pending, *vgive = &s_giveup;
    // lane 0 of a warp only; a lock that cannot be had (never seen) gives the segment up instead of hanging the GPU
    auto lock = [&]() {
        if (NW == 1) return;
        int spins = 0;
        while (atomicCAS(&s_lock, 0, 1) != 0) {
            __nanosleep(20);
            if (++spins > (1 << 24)) { *vgive = 1; break; }
        }
        __threadfence_block();
    };
    auto unlock = [&]() { if (NW == 1) return; __threadfence_block(); atomicExch(&s_lock, 0); };
    auto push = [&](int32_t lo_, int32_t hi_, int32_t it_) {   // lane 0 only, after a __syncwarp() behind the warp's last
        lock();                                                // write to x: the range's elements travel with the entry
        const int32_t t = *vsp;
        if (t < ICL_XS_STACK) {
            st_lo[t] = lo_; st_hi[t] = hi_; st_it[t] = it_;
            __threadfence_block();
            *vsp = t + 1;
            atomicAdd(&s_pending, 1);
        } else *vgive = 1;                                     // stack full (never seen): the host ranks this segment
        unlock();
    };
    int nsmall = 0;
    bool give_up = false;
    auto flush_small = [&]() {                                // std::sort sub-ranges, one per lane
        __syncwarp();
        int ok = 1;
        if (lane < nsmall) ok = xs_std_seq(v, x, sm_lo[lane], sm_hi[lane], sm_d[lane]) ? 1 : 0;
        if (__any_sync(FULL, !ok)) { give_up = true; if (lane == 0) *vgive = 1; }
        nsmall = 0;
        __syncwarp();
    };
    for (;;) {
        // ---- pop a range (lane 0), or leave when the stack is empty and no warp is still producing; idle warps look at
        // the stack pointer without taking the lock
        int32_t left = 0, right = 0, iters = 0, got = 0;
        if (lane == 0) {
            for (int spins = 0;; ++spins) {
                if (*vgive) break;
                if (spins > (1 << 21)) { *vgive = 1; break; }  // ~1 s without work while work is pending: give up
                if (*vsp > 0) {
                    lock();
                    const int32_t t = *vsp - 1;
                    if (t >= 0) {
                        left = st_lo[t]; right = st_hi[t]; iters = st_it[t];
                        *vsp = t;
                        got = 1;
                    }
                    unlock();
                    if (got) break;
                    continue;
                }
                if (*vpend == 0) break;
                __nanosleep(400);
            }
        }
        got = __shfl_sync(FULL, got, 0);
        if (!got) break;
        left = __shfl_sync(FULL, left, 0); right = __shfl_sync(FULL, right, 0); iters = __shfl_sync(FULL, iters, 0);
        __syncwarp();                                         // the other lanes read x behind lane 0's acquire
        // `done`: this popped range has been dealt with (its pieces are on the stack or in this warp's small list)
        auto done = [&]() { if (lane == 0) { __threadfence_block(); atomicSub(&s_pending, 1); } };
        if (iters < 0) {
            // ------------------------------------------------ std::sort range [f, l), depth budget -iters - 1
            int32_t f = left, l = right;
            int depth = -iters - 1;
            while (l - f > A.lane_max) {
                if (depth == 0) { give_up = true; if (lane == 0) *vgive = 1; break; }
                --depth;
                int32_t pv = 0;
                if (lane == 0) {                             // __move_median_to_first(f, f + 1, mid, l - 1)
                    const int32_t mid = f + ((l - f) >> 1), a = f + 1, b = mid, c = l - 1;
                    const int32_t ka = v[x[a]], kb = v[x[b]], kc = v[x[c]];
                    int32_t r;
                    if (ka < kb) { if (kb < kc) r = b; else if (ka < kc) r = c; else r = a; }
                    else if (ka < kc) r = a;
                    else if (kb < kc) r = c;
                    else r = b;
                    const IdxT t = x[f]; x[f] = x[r]; x[r] = t;
                    pv = v[x[f]];
                }
                pv = __shfl_sync(FULL, pv, 0);
                __syncwarp();
                // __unguarded_partition(f + 1, l, pivot at f): left stops = positions in [f + 1, l) with key >= pivot
                // (ascending), right stops = positions in [f + 1, l) with key <= pivot (taken back to front); pairs are
                // swapped while the left stop lies left of the right stop; the scan ends at the next left stop or at the
                // slot the last swap filled, whichever comes first
                int32_t *Ls = Lbase + f, *Rs = Rbase + f;
                int32_t nL = 0, nR = 0;
                bool alleq = true;
                for (int32_t b0 = f + 1; b0 < l; b0 += 32) {
                    const int32_t q = b0 + lane;
                    const int32_t kq = q < l ? v[x[q]] : pv;
                    const bool fl = q < l && !(kq < pv), fr = q < l && !(pv < kq);
                    const uint32_t ml = __ballot_sync(FULL, fl), mr = __ballot_sync(FULL, fr);
                    const uint32_t below = (1u << lane) - 1u;
                    alleq = alleq && kq == pv;
                    if (fl) Ls[nL + __popc(ml & below)] = q;
                    if (fr) Rs[nR + __popc(mr & below)] = q;
                    nL += __popc(ml);
                    nR += __popc(mr);
                }
                __syncwarp();
                if (__all_sync(FULL, alleq) && depth + 1 >= xs_std_equal_levels(l - f)) {
                    // equal keys throughout: the rest of std::sort on this range is a fixed permutation (xs_std_equal_pos);
                    // undo this level's median swap, apply all levels at once through the scratch slice
                    if (lane == 0) { const int32_t mid = f + ((l - f) >> 1); const IdxT t = x[f]; x[f] = x[mid]; x[mid] = t; }
                    __syncwarp();
                    for (int32_t q = f + lane; q < l; q += 32) Ls[xs_std_equal_pos(f, l, q) - f] = (int32_t)x[q];
                    __syncwarp();
                    for (int32_t q = f + lane; q < l; q += 32) x[q] = (IdxT)Ls[q - f];
                    __syncwarp();
                    l = f;                                   // nothing left of this range
                    break;
                }
                const int32_t kmax2 = nL < nR ? nL : nR;
                int32_t K = 0;
                for (int32_t b0 = 0; b0 < kmax2; b0 += 32) {
                    const int32_t k = b0 + lane;
                    const bool g = k < kmax2 && Ls[k] < Rs[nR - 1 - k];
                    const uint32_t m = __ballot_sync(FULL, g);
                    K += __popc(m);
                    if (m != FULL) break;
                }
                for (int32_t k = lane; k < K; k += 32) {
                    const int32_t a = Ls[k], b = Rs[nR - 1 - k];
                    const IdxT t = x[a]; x[a] = x[b]; x[b] = t;
                }
                __syncwarp();
                // the final left scan: next untouched left stop, or the slot the last swap filled; without either, l
                int32_t cut = K < nL ? Ls[K] : l;
                if (K >= 1 && Rs[nR - K] < cut) cut = Rs[nR - K];
                // right part [cut, l) onto the stack for any warp, go on with [f, cut)
                if (lane == 0) push(cut, l, -depth - 1);
                l = cut;
                __syncwarp();
            }
            if (!give_up && l - f > 1) {
                if (lane == 0) { sm_lo[nsmall] = f; sm_hi[nsmall] = l; sm_d[nsmall] = depth; }
                nsmall++;
                if (nsmall == 32) flush_small();
            }
            done();
            if (give_up) break;
            continue;
        }
        if (iters == 0) {                                     // argsort_ out of budget: std::sort of the range
            const int32_t cnt = right + 1 - left;
            if (lane == 0) push(left, right + 1, -(2 * (31 - __clz((unsigned)cnt))) - 1);
            __syncwarp();
            done();
            continue;
        }
        if (right + 1 - left <= 256) {
            xs_network(v, x, left, right + 1 - left, nkv, lane);
            done();
            continue;
        }
        // ---- pivot: 5th smallest of 8 samples (AVX-512), 3rd smallest of 4 (AVX2)
        auto median = [&](int32_t mine) {
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int32_t o = __shfl_sync(FULL, mine, j);
                rank += j < W && ((o < mine) || (o == mine && j < lane));
            }
            const uint32_t who = __ballot_sync(FULL, lane < W && rank == (W >> 1));
            return __shfl_sync(FULL, mine, __ffs(who) - 1);
        };
        const int32_t ssize = (right - left) >> wsh;
        const int32_t mine0 = lane < W ? v[x[left + (lane + 1) * ssize]] : 0x7fffffff;
        int32_t pivot = median(mine0);
        // ---- passes that cannot make progress (pivot == smallest key of the range), applied at once
        if (__reduce_min_sync(FULL, mine0) == pivot) {
            int32_t rmin = 0x7fffffff, rmax = (int32_t)0x80000000;
            for (int32_t i = left + lane; i <= right; i += 32) {
                const int32_t kv = v[x[i]];
                rmin = kv < rmin ? kv : rmin;
                rmax = kv > rmax ? kv : rmax;
            }
            rmin = __reduce_min_sync(FULL, rmin);
            rmax = __reduce_max_sync(FULL, rmax);
            if (pivot == rmin) {
                XsPerm P;
                P.left = left; P.R = right + 1;
                P.pre = (P.R - left) & (G - 1);
                P.rend = P.R - P.pre;
                P.M = (P.rend - left) >> (wsh + 2);
                int32_t q = lane < W ? left + (lane + 1) * ssize : left;
                int t = 0, state;
                for (;;) {
                    ++t;
                    if (pivot == rmax) { state = 0; break; }              // all keys equal: neither side is visited
                    if (iters - t == 0) { state = 1; break; }             // out of budget: std::sort of the range
                    q = xs_inv<W>(P, q);                                  // what sits at the sample positions now
                    pivot = median(lane < W ? v[x[q]] : 0x7fffffff);
                    if (pivot != rmin) { state = 2; break; }
                }
                int32_t *tmp = Lbase + left;                              // out of place through the scratch slice
                for (int32_t i = left + lane; i < P.R; i += 32) tmp[xs_fwd_t<W>(P, i, t) - left] = (int32_t)x[i];
                __syncwarp();
                for (int32_t i = left + lane; i < P.R; i += 32) x[i] = (IdxT)tmp[i - left];
                __syncwarp();
                iters -= t;
                if (state == 0) { done(); continue; }
                if (state == 1) {
                    const int32_t cnt = right + 1 - left;
                    if (lane == 0) push(left, right + 1, -(2 * (31 - __clz((unsigned)cnt))) - 1);
                    __syncwarp();
                    done();
                    continue;
                }
            }
        }
        // ---- partition [left, rend)
        const int32_t range_left = left, range_right = right;
        int32_t rend = right + 1;
        int32_t kmin = 0x7fffffff, kmax = (int32_t)0x80000000;
        {
            int32_t pre = (rend - left) & (G - 1);
            if (lane == 0) {
                for (; pre > 0; --pre) {
                    const int32_t kv = v[x[left]];
                    kmin = kv < kmin ? kv : kmin;
                    kmax = kv > kmax ? kv : kmax;
                    if (!(kv < pivot)) { --rend; const IdxT t = x[left]; x[left] = x[rend]; x[rend] = t; }
                    else ++left;
                }
            }
            left = __shfl_sync(FULL, left, 0);
            rend = __shfl_sync(FULL, rend, 0);
            __syncwarp();
        }
        int32_t l_store = left;
        if (left != rend) {
            const bool act = lane < G;                        // AVX2: 16 lanes carry a group
            const int32_t aL = act ? (int32_t)x[left + lane] : 0, kL = act ? v[aL] : 0;
            const int32_t aR = act ? (int32_t)x[rend - G + lane] : 0, kR = act ? v[aR] : 0;
            int32_t r_store = rend - W;
            int32_t lo = left + G, hi = rend - G;
            const int vec = lane >> wsh, lW = lane & (W - 1);
            const uint32_t maskW = (1u << W) - 1u, belowW = (1u << lW) - 1u;
            constexpr bool rev = W == 4;
            auto put = [&](int32_t a, int32_t kv) {          // one group of 4 vectors, vector by vector
                const bool isge = act && kv >= pivot;
                const uint32_t m = __ballot_sync(FULL, isge);
                const int geb = __popc(m & ((1u << (W * vec)) - 1u));   // keys >= pivot in the vectors before mine
                const int32_t ls = l_store + W * vec - geb, rs = r_store - geb;
                const uint32_t mv = (m >> (W * vec)) & maskW;
                const int ge = __popc(mv);
                const int rk = __popc(mv & belowW);
                const int32_t dst = isge ? (rev ? rs + W - 1 - rk : rs + W - ge + rk) : ls + __popc(~mv & belowW);
                __syncwarp();
                if (act) {
                    x[dst] = (IdxT)a;
                    kmin = kv < kmin ? kv : kmin;
                    kmax = kv > kmax ? kv : kmax;
                }
                const int tot = __popc(m);
                l_store += G - tot;
                r_store -= tot;
                __syncwarp();
            };
            // the next group of either side is fetched before the rule picks one: a pass only writes positions it has
            // already read, so reading early returns the same values and takes two shared-memory round trips off the chain
            int32_t nLa = 0, nLk = 0, nRa = 0, nRk = 0;
            auto fetch = [&](int32_t b0, int32_t &a, int32_t &k) { a = act ? (int32_t)x[b0 + lane] : 0; k = act ? v[a] : 0; };
            if (lo < hi) { fetch(lo, nLa, nLk); fetch(hi - G, nRa, nRk); }
            while (hi != lo) {
                int32_t a, k;
                if ((r_store + W) - hi < lo - l_store) { hi -= G; a = nRa; k = nRk; if (lo < hi) fetch(hi - G, nRa, nRk); }
                else { a = nLa; k = nLk; lo += G; if (lo < hi) fetch(lo, nLa, nLk); }
                put(a, k);
            }
            put(aL, kL);
            put(aR, kR);
        }
        kmin = __reduce_min_sync(FULL, kmin);
        kmax = __reduce_max_sync(FULL, kmax);
        const int32_t pi = l_store;
        // both sides with iters - 1 (order is free)
        if (lane == 0) {
            if (pivot != kmax) push(pi, range_right, iters - 1);
            if (pivot != kmin) push(range_left, pi - 1, iters - 1);
        }
        __syncwarp();
        done();
    }
    {
        const int gv = __shfl_sync(FULL, lane == 0 ? *vgive : 0, 0);
        if (nsmall && !gv) flush_small();
    }
    __syncthreads();
    if (threadIdx.x == 0 && *vgive) A.host_flag[it] = 1;
    if (fits) {
        int32_t *out = A.idx + base;
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) out[i] = (int32_t)x[i];
    }
}

template <class IdxT, int NW, int W>
static cudaError_t xss_launch(const XsArgs &A, unsigned grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_xss_argsort<IdxT, NW, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, ICL_XS_SMEM_MAX);
    if (e != cudaSuccess) return e;
    k_xss_argsort<IdxT, NW, W><<<grid, 32 * NW, smem, s>>>(A);
    return cudaGetLastError();
}
template <class IdxT, int W>
static cudaError_t xss_launch_w(int warps, const XsArgs &A, unsigned grid, size_t smem, cudaStream_t s) {
    switch (warps) {
    case 1: return xss_launch<IdxT, 1, W>(A, grid, smem, s);
    case 2: return xss_launch<IdxT, 2, W>(A, grid, smem, s);
    case 4: return xss_launch<IdxT, 4, W>(A, grid, smem, s);
    default: return xss_launch<IdxT, 8, W>(A, grid, smem, s);
    }
}
template <class IdxT>
static cudaError_t xss_launch_v(int lanes, int warps, const XsArgs &A, unsigned grid, size_t smem, cudaStream_t s) {
    return lanes == 4 ? xss_launch_w<IdxT, 4>(warps, A, grid, smem, s) : xss_launch_w<IdxT, 8>(warps, A, grid, smem, s);
}
#endif

#ifndef ICL_EMU
// numpy's argsort (flavour `mode`) of segments of `key` into `idx`.  off / order: device arrays (segment starts, segments
// longest first); lens[j] = length of segment order[j] (host copy, non-increasing).  Fails with ICL_ERR_STATE when a
// segment needs the host rank (std::sort's heap-sort branch).
static int rank_segments(icl_ctx *c, int32_t mode, const int32_t *key, int32_t *idx, const int64_t *off, const int32_t *order,
                         const std::vector<int64_t> &lens, bool defer = false) {
    Status &st = c->st;
    cudaStream_t s = c->stream;
    const int32_t nseg = (int32_t)lens.size();
    if (nseg == 0) return 0;
    int64_t total = 0;
    for (int64_t L : lens) total += L;
    if (c->npscr.reserve(st, 2 * (size_t)total + 64)) return st.code;
    const int64_t longest = lens[0];
    if (mode == ICL_RANK_NUMPY_SCALAR) {
        const bool narrow = longest <= 16000;                    // 16-bit indices: 6 B per element, everything on chip
        const int32_t cap = narrow ? (int32_t)((longest + 31) & ~31ll) : (int32_t)std::min<int64_t>((longest + 31) & ~31ll, 12288);
        const size_t smem = (size_t)cap * (narrow ? 6 : 8);
        NpSortArgs A{key, idx, c->npscr.p, off, order, cap};
        if (narrow) {
            ICL_CUDA(st, cudaFuncSetAttribute(k_np_argsort<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            k_np_argsort<uint16_t><<<(unsigned)nseg, 32 * ICL_NPS_WARPS, smem, s>>>(A);
        } else {
            ICL_CUDA(st, cudaFuncSetAttribute(k_np_argsort<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            k_np_argsort<int32_t><<<(unsigned)nseg, 32 * ICL_NPS_WARPS, smem, s>>>(A);
        }
        ICL_COUNT_KERNEL();
        ICL_CUDA(st, cudaGetLastError());
        return 0;
    }
    if (c->host_flag.reserve(st, (size_t)nseg + 1)) return st.code;
    ICL_CUDA(st, dev_memset(c->host_flag.p, 0, ((size_t)nseg + 1) * sizeof(int32_t), s));
    static const int xs_warps = [] { const char *e = getenv("ICL_B200_XS_WARPS"); return e ? atoi(e) : ICL_XS_WARPS_DEFAULT; }();
    static const int xs_lane = [] { const char *e = getenv("ICL_B200_XS_LANE"); const int q = e ? atoi(e) : ICL_XS_LANE; return q < 16 ? 16 : q > 160 ? 160 : q; }();
    static const int xs_coop = [] { const char *e = getenv("ICL_B200_XS_COOP"); return e ? atoi(e) : 1024; }();   // 0x7fffffff = off
    if (c->xs_chunk.reserve(st, 2 * ((size_t)(total >> 5) + 3 * (size_t)nseg) + 64)) return st.code;
    // Segments come longest first, so the size classes are runs of the order; one launch each:
    //   <= 16 000 keys   keys + 16-bit indices in shared memory (96 KB: several CTAs per SM)
    //   <= 34 000        the same with up to 204 KB (one CTA per SM)
    //   <= 65 535        16-bit indices in shared memory, keys through L1
    //   longer           32-bit indices; both in shared memory up to 12 288 keys, else in global memory
    auto cls_of = [](int64_t L) { return L > 65535 ? 0 : L > 34000 ? 1 : L > 16000 ? 2 : 3; };
    for (int32_t j0 = 0; j0 < nseg;) {
        const int64_t L = lens[j0];
        const int cls = cls_of(L);
        int32_t j1 = j0;
        while (j1 < nseg && cls_of(lens[j1]) == cls) j1++;
        const int64_t Lr = (L + 31) & ~31ll;
        const int32_t ccap = cls == 0 ? (int32_t)std::min<int64_t>(Lr, 12288) : (int32_t)Lr;
        const size_t csmem = (size_t)ccap * (cls == 0 ? 8 : cls == 1 ? 2 : 6);
        XsArgs A{key, idx, off, order, c->host_flag.p, c->npscr.p, ccap,
                 mode == ICL_RANK_NUMPY_AVX2 ? 4 : 8, j0, cls == 1 ? 0 : 1, xs_lane, getenv("ICL_B200_XS_FORCE_GIVEUP") ? 1 : 0,
                 xs_coop, c->xs_chunk.p};
        const unsigned grid = (unsigned)(j1 - j0);
        if (cls == 0) ICL_CUDA(st, xss_launch_v<int32_t>(A.W, xs_warps, A, grid, csmem, s));
        else ICL_CUDA(st, xss_launch_v<uint16_t>(A.W, xs_warps, A, grid, csmem, s));
        ICL_COUNT_KERNEL();
        j0 = j1;
    }
    ICL_CUDA(st, cudaGetLastError());
    // segments the kernel does not cover (std::sort's heap-sort branch) must be ranked by numpy itself: every such
    // segment raises host_flag[nseg]; `defer` leaves the look at it to the caller's next synchronisation (icl_extract)
    {
        int32_t *hf = c->host_flag.p;
        ICL_CUDA(st, foreach(s, (int64_t)nseg, ICL_LAMBDA(int64_t j) { if (hf[j]) atomic_or_i32(&hf[nseg], 1); }));
    }
    if (defer) { c->rank_flag_pending = true; c->rank_flag_slot = nseg; return 0; }
    int32_t any = 0;
    ICL_CUDA(st, d2h(&any, c->host_flag.p + nseg, sizeof(int32_t), s));
    ICL_CUDA(st, dev_sync(s));
    if (any) return fail(c, ICL_ERR_STATE, "icl_rank_device: a segment needs the host rank (std::sort depth limit)");
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// ABI: icl_rank_device — the extraction order of every (pair, direction) item computed on the device.
// ICL_RANK_NUMPY_AVX512 / _AVX2 / _SCALAR: numpy's own argsort of that flavour, reproduced exactly (the reference's tie
// order wherever numpy runs that flavour).  ICL_RANK_STABLE: NON-REFERENCE tie order from a STABLE device sort — what
// pandas' nargsort(ascending=False) would return if numpy's argsort were stable (reverse, stable ascending argsort,
// reverse): identical to the reference wherever scores are distinct; tied scores come in reversed (loc1, loc2) order
// instead of numpy's CPU-specific permutation.
// ------------------------------------------------------------------------------------------------
extern "C" int icl_rank_device(icl_ctx *c, int32_t mode) {
    if (!c) return ICL_ERR_ARG;
    if (!c->dp_done) return fail(c, ICL_ERR_STATE, "icl_rank_device: icl_dp has not succeeded");
    if (mode != ICL_RANK_STABLE && mode != ICL_RANK_NUMPY_SCALAR && mode != ICL_RANK_NUMPY_AVX512 && mode != ICL_RANK_NUMPY_AVX2)
        return fail(c, ICL_ERR_ARG, "icl_rank_device: unknown mode");
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    cudaStream_t s = c->stream;
    const int64_t n = c->n, V = 2 * n;
    const int32_t P = c->P;
    LaunchScope scope(c, 1);
    c->rank_flag_pending = false;
    c->timers[ICL_T_D2H].start(s);
    if (c->order_in.reserve(st, V) || c->key_a.reserve(st, V) || c->key_b.reserve(st, V) || c->idx_a.reserve(st, V) ||
        c->tree_v.reserve(st, V))
        return st.code;
    if (n > 0 && mode != ICL_RANK_STABLE) {
        // the REVERSED score array of every (pair, direction) item, as in the score export of icl_dp
        if (c->score32.reserve(st, V) || c->npscr.reserve(st, 2 * V)) return st.code;
        const int32_t *S = c->S.p, *mmap = c->mmap.p, *ps = c->pair_s.p;
        const int64_t *off = c->d_pair_off.p;
        int32_t *o = c->score32.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            int32_t p = ps[i];
            int64_t src = off[p] + (off[p + 1] - 1 - i);
            o[i] = S[src];
            o[n + i] = S[mmap[src]];
        }));
#ifndef ICL_EMU
        std::vector<int64_t> lens(2 * (size_t)P);
        for (int32_t j = 0; j < 2 * P; j++) { const int32_t it = c->h_item_order[j]; lens[j] = c->h_seg_off[it + 1] - c->h_seg_off[it]; }
        if (rank_segments(c, mode, c->score32.p, c->order_in.p, c->d_seg_off.p, c->d_item_order.p, lens, true)) return st.code;
#else
        return fail(c, ICL_ERR_STATE, "icl_rank_device: the numpy-scalar mode exists in the CUDA build only");
#endif
    } else if (n > 0) {
        // element i of the first half = position j of pair p's REVERSED segment (as in the score export of icl_dp);
        // key = item | score, value = j; a stable sort leaves every item's ascending argsort in place
        const int32_t *S = c->S.p, *mmap = c->mmap.p, *ps = c->pair_s.p;
        const int64_t *off = c->d_pair_off.p;
        uint64_t *key = c->key_a.p;
        uint32_t *val = c->idx_a.p;
        ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            const int32_t p = ps[i];
            const int64_t src = off[p] + (off[p + 1] - 1 - i);
            const uint32_t j = (uint32_t)(i - off[p]);
            key[i] = ((uint64_t)(uint32_t)p << 32) | (uint32_t)S[src];
            key[n + i] = ((uint64_t)(uint32_t)(P + p) << 32) | (uint32_t)S[mmap[src]];
            val[i] = j;
            val[n + i] = j;
        }));
        if (radix_sort_pairs(st, c->tmp, c->key_a.p, c->key_b.p, c->idx_a.p, c->tree_v.p, V, 32 + bits_for(2 * (int64_t)P), s)) return st.code;
        ICL_CUDA(st, d2d(c->order_in.p, c->tree_v.p, (size_t)V * sizeof(int32_t), s));
    }
    c->timers[ICL_T_D2H].stop(s);
    c->order_on_device = true;
    c->keys_on_device = mode != ICL_RANK_STABLE;
    return ICL_OK;
}

// Probe / test: the rank kernels on caller-supplied key segments (nothing of the upload is touched).
extern "C" int icl_debug_rank_keys(icl_ctx *c, int32_t mode, int32_t n_segments, const int64_t *seg_off, const int32_t *keys,
                                   int32_t *argsort_out) {
    if (!c) return ICL_ERR_ARG;
    if (n_segments < 0 || !seg_off || (n_segments > 0 && seg_off[n_segments] > 0 && (!keys || !argsort_out)))
        return fail(c, ICL_ERR_ARG, "icl_debug_rank_keys: bad argument");
    if (mode != ICL_RANK_NUMPY_SCALAR && mode != ICL_RANK_NUMPY_AVX512 && mode != ICL_RANK_NUMPY_AVX2)
        return fail(c, ICL_ERR_ARG, "icl_debug_rank_keys: mode must be one of the numpy flavours");
#ifndef ICL_EMU
    Status &st = c->st;
    st.code = 0;
    ICL_CUDA(st, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (seg_off[0] != 0) return fail(c, ICL_ERR_ARG, "icl_debug_rank_keys: seg_off[0] must be 0");
    for (int32_t j = 0; j < n_segments; j++)
        if (seg_off[j + 1] < seg_off[j] || seg_off[j + 1] - seg_off[j] >= (1 << 24))
            return fail(c, ICL_ERR_ARG, "icl_debug_rank_keys: segment offsets must ascend, segments < 2^24 keys");
    const int64_t total = n_segments ? seg_off[n_segments] : 0;
    if (total == 0) return ICL_OK;
    std::vector<int32_t> order((size_t)n_segments);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return seg_off[a + 1] - seg_off[a] > seg_off[b + 1] - seg_off[b]; });
    std::vector<int64_t> lens((size_t)n_segments);
    for (int32_t j = 0; j < n_segments; j++) lens[j] = seg_off[order[j] + 1] - seg_off[order[j]];
    Buf<int32_t> dk, di, dord;
    Buf<int64_t> doff;
    int rc = 0;
    if (dk.reserve(st, total) || di.reserve(st, total) || dord.reserve(st, n_segments) || doff.reserve(st, (size_t)n_segments + 1)) rc = st.code;
    if (!rc) {
        cudaError_t e = h2d(dk.p, keys, (size_t)total * 4, s);
        if (e == cudaSuccess) e = h2d(dord.p, order.data(), (size_t)n_segments * 4, s);
        if (e == cudaSuccess) e = h2d(doff.p, seg_off, ((size_t)n_segments + 1) * 8, s);
        if (e != cudaSuccess) rc = fail(c, ICL_ERR_CUDA, cudaGetErrorString(e));
    }
    if (!rc) rc = rank_segments(c, mode, dk.p, di.p, doff.p, dord.p, lens);
    if (!rc) {
        cudaError_t e = d2h(argsort_out, di.p, (size_t)total * 4, s);
        if (e == cudaSuccess) e = dev_sync(s);
        if (e != cudaSuccess) rc = fail(c, ICL_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaStreamSynchronize(s);
    dk.release(); di.release(); dord.release(); doff.release();
    return rc;
#else
    return fail(c, ICL_ERR_STATE, "icl_debug_rank_keys: the numpy rank kernels exist in the CUDA build only");
#endif
}

// Test / verification: the key arrays the last icl_rank_device (numpy flavours) sorted — every (pair, direction) item's
// scores in REVERSED point order, forward items then anti-diagonal ones (2 n values; what icl_dp would have exported).
extern "C" int icl_debug_keys(icl_ctx *c, int32_t *keys) {
    if (!c || !keys) return ICL_ERR_ARG;
    if (!c->order_on_device || !c->keys_on_device) return fail(c, ICL_ERR_STATE, "icl_debug_keys: icl_rank_device (numpy flavour) has not succeeded");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    if (c->n > 0) {
        ICL_CUDA(st, d2h(keys, c->score32.p, (size_t)2 * c->n * sizeof(int32_t), c->stream));
        ICL_CUDA(st, dev_sync(c->stream));
    }
    return ICL_OK;
}

extern "C" int icl_debug_order(icl_ctx *c, int32_t *argsort) {
    if (!c || !argsort) return ICL_ERR_ARG;
    if (!c->order_on_device) return fail(c, ICL_ERR_STATE, "icl_debug_order: icl_rank_device has not succeeded");
    Status &st = c->st;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    if (c->n > 0) {
        ICL_CUDA(st, d2h(argsort, c->order_in.p, (size_t)2 * c->n * sizeof(int32_t), c->stream));
        ICL_CUDA(st, dev_sync(c->stream));
    }
    return ICL_OK;
}

extern "C" int icl_extract(icl_ctx *c, const int32_t *argsort_fwd, const int32_t *argsort_rev, int32_t min_members,
                           icl_blocks *out) {
    return icl_extract_fields(c, argsort_fwd, argsort_rev, min_members, ICL_FIELDS_ALL, out);
}

extern "C" int icl_extract_fields(icl_ctx *c, const int32_t *argsort_fwd, const int32_t *argsort_rev, int32_t min_members,
                                  int32_t fields, icl_blocks *out) {
    if (!c) return ICL_ERR_ARG;
    if (fields != ICL_FIELDS_ALL && fields != ICL_FIELDS_WRITER) return fail(c, ICL_ERR_ARG, "icl_extract_fields: unknown field set");
    c->out_fields = fields;
    if (!c->dp_done) return fail(c, ICL_ERR_STATE, "icl_extract: icl_dp has not succeeded");
    const bool dev_order = !argsort_fwd && !argsort_rev && c->order_on_device;       // icl_rank_device ran
    if (!out || (c->n > 0 && !dev_order && (!argsort_fwd || !argsort_rev))) return fail(c, ICL_ERR_ARG, "icl_extract: null argument");
    memset(out, 0, sizeof(*out));
    Status &st = c->st;
    st.code = 0;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaSetDevice(c->device));
#endif
    cudaStream_t s = c->stream;
    const int64_t n = c->n, V = 2 * n;
    LaunchScope scope(c, 2);
    c->timers[ICL_T_EXTRACT].start(s);
    if (c->order_in.reserve(st, 2 * n) || c->rank2v.reserve(st, V)) return st.code;
    c->order_on_device_used = dev_order;
    if (n > 0) {
        if (dev_order) {
            // order_in already holds the device argsort
        } else if (argsort_rev == argsort_fwd + n) {
            ICL_CUDA(st, h2d(c->order_in.p, argsort_fwd, (size_t)2 * n * sizeof(int32_t), s));
        } else {
            ICL_CUDA(st, h2d(c->order_in.p, argsort_fwd, (size_t)n * sizeof(int32_t), s));
            ICL_CUDA(st, h2d(c->order_in.p + n, argsort_rev, (size_t)n * sizeof(int32_t), s));
        }
        // rank2v[V position] = point examined at that rank position of its item (a device order: in phase_extract's key pass)
        const int32_t *ps = c->pair_s.p, *mmap = c->mmap.p, *ord = c->order_in.p;
        const int64_t *off = c->d_pair_off.p;
        int32_t *rank2v = c->rank2v.p;
        if (!dev_order) ICL_CUDA(st, dev_memset(c->rank2v.p, 0xFF, (size_t)V * sizeof(int32_t), s));
        if (!dev_order) ICL_CUDA(st, foreach(s, n, ICL_LAMBDA(int64_t i) {
            int32_t p = ps[i];
            int64_t a = off[p], np_ = off[p + 1] - a, j = i - a;
            int32_t rf = ord[i], rr = ord[n + i];
            if (rf < 0 || rf >= np_ || rr < 0 || rr >= np_) return;               // stays -1: reported by phase_extract
            rank2v[a + (np_ - 1 - j)] = (int32_t)(a + (np_ - 1 - rf));
            rank2v[n + a + (np_ - 1 - j)] = mmap[a + (np_ - 1 - rr)];
        }));
    }
    c->min_members = min_members;
    if (phase_extract(c)) return st.code;
    // With the device's own order (nothing of the caller's to validate) the times pass does not wait for the host: its
    // inputs go into arrays sized by a bound, the kernel runs, and the host meanwhile receives the counts, sizes the output
    // staging and queues the output passes on the side stream.  Otherwise: counts first, then the same passes.
    cudaStream_t s2 = c->stream2 ? c->stream2 : s;
    const bool overlap = c->order_on_device_used && c->pvalue_mode != 2 && n > 0;
    if (!overlap && extract_counts(c)) return st.code;
    if (extract_block_slices(c, overlap ? extract_block_bound(c) : c->NB)) return st.code;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventRecord(c->ev_fork, s));              // what the output passes on the side stream wait for
#endif
    c->timers[ICL_T_EXTRACT].stop(s);
    if (overlap) {
        c->timers[ICL_T_PVALUE].start(s);
        c->NB = extract_block_bound(c);                        // (capacity of the per-block arrays; the kernels read the real lists)
        if (phase_pvalue(c)) return st.code;
        c->timers[ICL_T_PVALUE].stop(s);
        if (extract_counts(c)) return st.code;
    }
    const int64_t NBo = c->NBo, NMo = c->NMo;
    if (stage_layout(c, NBo, NMo)) return st.code;
    // the gene pairs of the output do not wait for the p values: gathered and copied on the side stream meanwhile
#ifndef ICL_EMU
    ICL_CUDA(st, cudaStreamWaitEvent(s2, c->ev_fork, 0));
#endif
    if (extract_output_records(c, s2)) return st.code;
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventRecord(c->ev_join3, s2));            // the records: the p values' pass below reads them
#endif
    if (phase_members(c, s2)) return st.code;
    {
        const size_t end = c->out_fields == ICL_FIELDS_WRITER ? c->stage_members_lean_end : c->stage_bytes;
        if (end > c->stage_members_at)
            ICL_CUDA(st, d2h(c->h_stage + c->stage_members_at, c->d_stage.p + c->stage_members_at, end - c->stage_members_at, s2));
        // ... and the per-block arrays the times pass does not fill (written by extract_output_records above)
        if (c->stage_praw_at > 0) ICL_CUDA(st, d2h(c->h_stage, c->d_stage.p, c->stage_praw_at, s2));
        if (c->out_fields != ICL_FIELDS_WRITER && c->stage_late_at > c->stage_blocks_lean_end)
            ICL_CUDA(st, d2h(c->h_stage + c->stage_blocks_lean_end, c->d_stage.p + c->stage_blocks_lean_end,
                             c->stage_late_at - c->stage_blocks_lean_end, s2));
    }
#ifndef ICL_EMU
    ICL_CUDA(st, cudaEventRecord(c->ev_join, s2));
#endif
    if (!overlap) {
        c->timers[ICL_T_PVALUE].start(s);
        if (phase_pvalue(c)) return st.code;
        c->timers[ICL_T_PVALUE].stop(s);
    }
    c->timers[ICL_T_GATHER].start(s);
#ifndef ICL_EMU
    ICL_CUDA(st, cudaStreamWaitEvent(s, c->ev_join3, 0));
#endif
    {
        // N, N1 and the p value (collinearity.py:152 in binary64, left to right) of the blocks the caller gets
        const uint32_t *o_src = c->o_src.p;
        const int32_t *b_N = c->b_N.p;
        const int64_t *b_N1 = c->b_N1.p;
        const int32_t g0 = c->prm.g0;
        const icl_blocks o = c->dv;
        ICL_CUDA(st, foreach(s, NBo, ICL_LAMBDA(int64_t g) {
            const uint32_t b = o_src[g];
            const int32_t N = b_N[b];
            const int64_t N1 = b_N1[b];
            o.N[g] = N;
            o.N1[g] = N1;
            o.p_raw[g] = icl_p_raw(o.score[g], o.m[g], g0, N1, N, o.L1[g], o.L2[g]);
        }));
    }
    // behind the times pass: p_raw (and N, N1)
    ICL_CUDA(st, d2h(c->h_stage + c->stage_praw_at, c->d_stage.p + c->stage_praw_at, c->stage_blocks_lean_end - c->stage_praw_at, s));
    if (c->out_fields != ICL_FIELDS_WRITER)
        ICL_CUDA(st, d2h(c->h_stage + c->stage_late_at, c->d_stage.p + c->stage_late_at, c->stage_members_at - c->stage_late_at, s));
#ifndef ICL_EMU
    ICL_CUDA(st, cudaStreamWaitEvent(s, c->ev_join, 0));
#endif
    c->timers[ICL_T_GATHER].stop(s);
    ICL_CUDA(st, dev_sync(s));
    *out = c->hv;
    if (c->out_fields == ICL_FIELDS_WRITER)                    // not copied: not there
        out->dir = out->L = out->N = out->L1 = out->L2 = out->member_idx = nullptr, out->N1 = nullptr;
    out->n_blocks = NBo;
    out->n_members = NMo;
    return ICL_OK;
}

extern "C" int icl_timings(icl_ctx *c, double *ms, int32_t n_slots) {
    if (!c || !ms) return ICL_ERR_ARG;
    int32_t k = n_slots < ICL_T_COUNT ? n_slots : ICL_T_COUNT;
    for (int32_t i = 0; i < k; i++) ms[i] = c->timers[i].ms();
    return k;
}

extern "C" int icl_cache_stats(int64_t *out, int32_t n_slots) {
    if (!out) return ICL_ERR_ARG;
    BlockCache &bc = BlockCache::get();
    const int64_t v[3] = {bc.hits, bc.misses, (int64_t)bc.cached_bytes()};
    int32_t k = n_slots < 3 ? n_slots : 3;
    for (int32_t i = 0; i < k; i++) out[i] = v[i];
    return k;
}

extern "C" void icl_cache_trim(void) {
    BlockCache &bc = BlockCache::get();
    std::lock_guard<std::mutex> lock(bc.mu);
    for (auto &kv : bc.pools) {
#ifndef ICL_EMU
        if (kv.first >= 0) cudaSetDevice(kv.first);
#endif
        bc.trim(kv.first);
    }
}

extern "C" int icl_counters(icl_ctx *c, int64_t *out, int32_t n_slots) {
    if (!c || !out) return ICL_ERR_ARG;
    int32_t k = n_slots < 8 ? n_slots : 8;
    for (int32_t i = 0; i < k; i++)
        out[i] = i == 1 ? c->n_own[0] : i == 2 ? c->n_own[2] : i < 4 ? c->counters[i] : i < 6 ? c->counters_ext[i - 4]
                 : i == 6 ? c->n_own[1] : c->n_cub[0] + c->n_cub[1] + c->n_cub[2];
    return k;
}
