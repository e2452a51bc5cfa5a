// icl_platform.h — thin device layer of the icl_b200 library.
//
// Real build (nvcc, sm_100a): CUDA runtime + CUB.
// ICL_EMU build (g++, test-suite only): the same kernel bodies run as host loops so that the
// host-side logic (layout, compaction, ABI) can be exercised on a machine without a GPU.  The
// emulation is never loaded by the wgdi_b200 package (see wgdi_b200/_lib.py).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>

#ifndef ICL_EMU
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#define ICL_HD __host__ __device__ __forceinline__
#define ICL_D __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define ICL_LDG(p) __ldg(p)          // read-only data through L1
#else
#define ICL_LDG(p) (*(p))
#endif
#else
#include <algorithm>
#include <chrono>
#include <numeric>
#include <vector>
#define ICL_HD inline
#define ICL_D inline
#define ICL_LDG(p) (*(p))
typedef int cudaError_t;
typedef void *cudaStream_t;
#define cudaSuccess 0
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
struct int2 { int32_t x, y; };
static inline int2 make_int2(int32_t x, int32_t y) { int2 r = {x, y}; return r; }
struct int4 { int32_t x, y, z, w; };
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { int4 r = {x, y, z, w}; return r; }
#endif

namespace icl {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
struct Status {
    int code = 0;
    std::string msg;
    bool ok() const { return code == 0; }
};

#ifndef ICL_EMU
#define ICL_CUDA(st, call)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (st).code = -2;                                                                        \
            (st).msg = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return (st).code;                                                                      \
        }                                                                                          \
    } while (0)
#else
#define ICL_CUDA(st, call) do { (void)(call); } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
// Raw allocations: device memory and page-locked host memory.
#ifndef ICL_EMU
inline cudaError_t raw_dev_malloc(void **p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t raw_dev_free(void *p) { return cudaFree(p); }
inline cudaError_t raw_pin_malloc(void **p, size_t bytes) { return cudaHostAlloc(p, bytes, cudaHostAllocDefault); }
inline cudaError_t raw_pin_free(void *p) { return cudaFreeHost(p); }
inline void raw_dev_zero(void *p, size_t bytes) { cudaMemsetAsync(p, 0, bytes, 0); cudaStreamSynchronize(0); }
inline void raw_device_idle() { cudaDeviceSynchronize(); }
inline int raw_current_device() { int d = 0; cudaGetDevice(&d); return d; }
inline void raw_set_device(int d) { cudaSetDevice(d); }
inline void raw_clear_error() { (void)cudaGetLastError(); }
inline cudaError_t h2d(void *d, const void *h, size_t b, cudaStream_t s) { return cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, s); }
inline cudaError_t d2h(void *h, const void *d, size_t b, cudaStream_t s) { return cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s); }
inline cudaError_t d2d(void *d, const void *s_, size_t b, cudaStream_t s) { return cudaMemcpyAsync(d, s_, b, cudaMemcpyDeviceToDevice, s); }
inline cudaError_t dev_memset(void *d, int v, size_t b, cudaStream_t s) { return cudaMemsetAsync(d, v, b, s); }
inline cudaError_t dev_sync(cudaStream_t s) { return cudaStreamSynchronize(s); }
#else
inline int raw_dev_malloc(void **p, size_t bytes) { *p = malloc(bytes); return *p ? 0 : 1; }
inline int raw_dev_free(void *p) { free(p); return 0; }
inline int raw_pin_malloc(void **p, size_t bytes) { *p = malloc(bytes); return *p ? 0 : 1; }
inline int raw_pin_free(void *p) { free(p); return 0; }
inline void raw_dev_zero(void *p, size_t bytes) { memset(p, 0, bytes); }
inline void raw_device_idle() {}
inline int raw_current_device() { return 0; }
inline void raw_set_device(int) {}
inline void raw_clear_error() {}
inline int h2d(void *d, const void *h, size_t b, cudaStream_t) { memcpy(d, h, b); return 0; }
inline int d2h(void *h, const void *d, size_t b, cudaStream_t) { memcpy(h, d, b); return 0; }
inline int d2d(void *d, const void *s_, size_t b, cudaStream_t) { memcpy(d, s_, b); return 0; }
inline int dev_memset(void *d, int v, size_t b, cudaStream_t) { memset(d, v, b); return 0; }
inline int dev_sync(cudaStream_t) { return 0; }
#endif

// Process-wide cache of released blocks.  A context holds ~90 device arrays sized by its comparison; cudaMalloc and
// cudaFree of gigabytes cost far more than the step that uses them (C4: 0.1-0.5 s to tear a context down, as much again
// to build the next one, for an 8 ms step), and the drop-in driver, the batch runner and the sharded runs create one
// context per comparison.  A released block is kept (per device / for page-locked memory) and handed to the next request
// of about its size; nothing goes back to the driver before the process ends, unless the cache outgrows its limit
// (ICL_B200_CACHE_MB, default 16384; 0 switches the cache off) or an allocation fails (everything cached is freed and
// the allocation tried again).
//  * release keeps cudaFree's guarantee: the device is idle before the block can be handed out again;
//  * a recycled device block is zeroed before reuse — strictly more than cudaMalloc promises.
struct BlockCache {
    struct Pool {
        std::multimap<size_t, void *> free_blocks;
        size_t cached = 0;
    };
    std::mutex mu;
    std::map<int, Pool> pools;                       // device id; -1 = page-locked host memory
    std::unordered_map<void *, std::pair<int, size_t>> live;
    size_t limit;
    int64_t hits = 0, misses = 0;
    BlockCache() {
        const char *e = getenv("ICL_B200_CACHE_MB");
        limit = (size_t)(e ? strtoull(e, nullptr, 10) : 16384ull) << 20;
    }
    static BlockCache &get() {
        static BlockCache *c = new BlockCache();     // never destroyed: the driver may be gone before static destructors run
        return *c;
    }
    static size_t round_up(size_t bytes) { return (std::max<size_t>(bytes, 16) + 511) & ~(size_t)511; }
    void trim(int pool) {                            // everything cached in one pool back to the driver (mu held)
        Pool &pl = pools[pool];
        for (auto &kv : pl.free_blocks) {
            if (pool < 0) raw_pin_free(kv.second); else raw_dev_free(kv.second);
        }
        pl.free_blocks.clear();
        pl.cached = 0;
    }
    cudaError_t alloc(void **p, size_t bytes, bool pinned) {
        bytes = round_up(bytes);
        const int pool = pinned ? -1 : raw_current_device();
        {
            size_t got = 0;
            {
                std::lock_guard<std::mutex> lock(mu);
                Pool &pl = pools[pool];
                auto it = pl.free_blocks.lower_bound(bytes);
                if (it != pl.free_blocks.end() && it->first <= bytes + bytes / 4 + 4096) {
                    *p = it->second;
                    got = it->first;
                    pl.cached -= got;
                    pl.free_blocks.erase(it);
                    live[*p] = std::make_pair(pool, got);
                    hits++;
                } else {
                    misses++;
                }
            }
            if (got) {                               // zeroed outside the lock: other threads keep allocating meanwhile
                if (!pinned) raw_dev_zero(*p, got);
                return cudaSuccess;
            }
        }
        cudaError_t e = pinned ? raw_pin_malloc(p, bytes) : raw_dev_malloc(p, bytes);
        if (e != cudaSuccess) {                      // make room: the cache gives everything back, then once more
            raw_clear_error();
            {
                std::lock_guard<std::mutex> lock(mu);
                trim(pool);
            }
            e = pinned ? raw_pin_malloc(p, bytes) : raw_dev_malloc(p, bytes);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lock(mu);
            live[*p] = std::make_pair(pool, bytes);
        }
        return e;
    }
    cudaError_t release(void *p) {
        if (!p) return cudaSuccess;
        int pool = 0;
        size_t bytes = 0;
        bool known = false;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = live.find(p);
            if (it != live.end()) { pool = it->second.first; bytes = it->second.second; known = true; live.erase(it); }
        }
        if (!known) return cudaSuccess;              // not ours (never happens: every block comes from alloc)
        if (limit == 0) return pool < 0 ? raw_pin_free(p) : raw_dev_free(p);
        if (pool >= 0) {                             // cudaFree's implicit guarantee, kept: nothing on the device uses the block
            const int cur = raw_current_device();
            if (cur != pool) raw_set_device(pool);
            raw_device_idle();
            if (cur != pool) raw_set_device(cur);
        } else {
            raw_device_idle();                       // page-locked: copies of the calling context's device have landed
        }
        std::lock_guard<std::mutex> lock(mu);
        Pool &pl = pools[pool];
        if (pl.cached + bytes > limit) trim(pool);
        if (bytes > limit) return pool < 0 ? raw_pin_free(p) : raw_dev_free(p);
        pl.free_blocks.emplace(bytes, p);
        pl.cached += bytes;
        return cudaSuccess;
    }
    size_t cached_bytes() {
        std::lock_guard<std::mutex> lock(mu);
        size_t t = 0;
        for (auto &kv : pools) t += kv.second.cached;
        return t;
    }
};

inline cudaError_t dev_malloc(void **p, size_t bytes) { return BlockCache::get().alloc(p, bytes, false); }
inline cudaError_t dev_free(void *p) { return BlockCache::get().release(p); }
inline cudaError_t pin_malloc(void **p, size_t bytes) { return BlockCache::get().alloc(p, bytes, true); }
inline cudaError_t pin_free(void *p) { return BlockCache::get().release(p); }

// grow-only device buffer
template <class T>
struct Buf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(Status &st, size_t n) {
        if (n <= cap && p) return 0;
        if (p) dev_free(p);
        p = nullptr;
        size_t want = n + n / 8 + 64;
        cap = 0;
        ICL_CUDA(st, dev_malloc((void **)&p, want * sizeof(T)));
        cap = want;
        return 0;
    }
    void release() { if (p) dev_free(p); p = nullptr; cap = 0; }
};

// ------------------------------------------------------------------------------------------------
// launch accounting: every kernel launch of the library and every CUB device-wide call on the calling thread
// (an ABI call runs on one host thread from start to finish: it reads the counters on entry and on exit)
// ------------------------------------------------------------------------------------------------
struct LaunchCount { int64_t own = 0, cub = 0; };
inline LaunchCount &launch_count() { static thread_local LaunchCount lc; return lc; }
#define ICL_COUNT_KERNEL() (++icl::launch_count().own)

// ------------------------------------------------------------------------------------------------
// kernel launch: grid-stride foreach over [0, n)
// ------------------------------------------------------------------------------------------------
#ifndef ICL_EMU
template <class F>
__global__ void __launch_bounds__(256) k_foreach(int64_t n, F f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i);
}
// one element per block-thread group of `tpb` threads: used for one-thread-per-item kernels where
// every item wants its own warp (no intra-warp divergence between unrelated items)
template <class F>
__global__ void k_foreach_warp(int64_t n, F f) {
    int64_t i = blockIdx.x;
    if (i < n && threadIdx.x == 0) f(i);
}
template <class F>
inline cudaError_t foreach(cudaStream_t s, int64_t n, F f) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    const int64_t maxb = 148 * 16;
    if (blocks > maxb) blocks = maxb;
    k_foreach<<<(unsigned)blocks, 256, 0, s>>>(n, f);
    ICL_COUNT_KERNEL();
    return cudaGetLastError();
}
template <class F>
inline cudaError_t foreach_warp(cudaStream_t s, int64_t n, F f) {
    if (n <= 0) return cudaSuccess;
    k_foreach_warp<<<(unsigned)n, 32, 0, s>>>(n, f);
    ICL_COUNT_KERNEL();
    return cudaGetLastError();
}
#else
template <class F> inline int foreach(cudaStream_t, int64_t n, F f) { if (n > 0) ICL_COUNT_KERNEL(); for (int64_t i = 0; i < n; i++) f(i); return 0; }
template <class F> inline int foreach_warp(cudaStream_t, int64_t n, F f) { if (n > 0) ICL_COUNT_KERNEL(); for (int64_t i = 0; i < n; i++) f(i); return 0; }
#endif

ICL_HD void atomic_max_i32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
ICL_HD void atomic_add_i32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
ICL_HD int32_t atomic_fetch_add_i32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const int32_t old = *p;
    *p += v;
    return old;
#endif
}
ICL_HD void atomic_add_i64(int64_t *p, int64_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd((unsigned long long *)p, (unsigned long long)v);
#else
    *p += v;
#endif
}
ICL_HD void atomic_or_i32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// ------------------------------------------------------------------------------------------------
// sort / scan (CUB on the device, std:: in the emulation)
// ------------------------------------------------------------------------------------------------
#ifndef ICL_EMU
inline int radix_sort_pairs(Status &st, Buf<uint8_t> &tmp, const uint64_t *kin, uint64_t *kout,
                            const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit, cudaStream_t s) {
    size_t bytes = 0;
    ICL_CUDA(st, cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    ++launch_count().cub;
    return 0;
}
inline int radix_sort_pairs32(Status &st, Buf<uint8_t> &tmp, const uint32_t *kin, uint32_t *kout,
                              const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit, cudaStream_t s) {
    size_t bytes = 0;
    if (end_bit > 32) end_bit = 32;
    ICL_CUDA(st, cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    ++launch_count().cub;
    return 0;
}
inline int radix_sort_keys(Status &st, Buf<uint8_t> &tmp, const uint64_t *kin, uint64_t *kout, int64_t n, int end_bit,
                           cudaStream_t s) {
    size_t bytes = 0;
    if (end_bit > 64) end_bit = 64;
    ICL_CUDA(st, cub::DeviceRadixSort::SortKeys(nullptr, bytes, kin, kout, (int)n, 0, end_bit, s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceRadixSort::SortKeys(tmp.p, bytes, kin, kout, (int)n, 0, end_bit, s));
    ++launch_count().cub;
    return 0;
}
// out[0..n] = exclusive prefix sum of in[0..n) with the total in out[n]; `in` must have n+1 slots
// (slot n is overwritten with 0 first)
inline int exclusive_sum_u32(Status &st, Buf<uint8_t> &tmp, uint32_t *in, uint32_t *out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    ICL_CUDA(st, dev_memset(in + n, 0, sizeof(uint32_t), s));
    ICL_CUDA(st, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)(n + 1), s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)(n + 1), s));
    ++launch_count().cub;
    return 0;
}
struct MaxI32 {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};
// the same for 64-bit words: two 32-bit counters travel in one word when neither sum reaches 2^32 (a struct of three
// counters goes through CUB's non-primitive tile states: 80 us for 3.8 M elements against 13-16 us for a primitive type)
inline int exclusive_sum_u64(Status &st, Buf<uint8_t> &tmp, uint64_t *in, uint64_t *out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    ICL_CUDA(st, dev_memset(in + n, 0, sizeof(uint64_t), s));
    ICL_CUDA(st, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)(n + 1), s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)(n + 1), s));
    ++launch_count().cub;
    return 0;
}
// in place: a[i] = max(a[0..i])
inline int inclusive_max_i32(Status &st, Buf<uint8_t> &tmp, int32_t *a, int64_t n, cudaStream_t s) {
    if (n <= 0) return 0;
    size_t bytes = 0;
    ICL_CUDA(st, cub::DeviceScan::InclusiveScan(nullptr, bytes, a, a, MaxI32(), (int)n, s));
    if (tmp.reserve(st, bytes)) return st.code;
    ICL_CUDA(st, cub::DeviceScan::InclusiveScan(tmp.p, bytes, a, a, MaxI32(), (int)n, s));
    ++launch_count().cub;
    return 0;
}
struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool armed = false;
    void start(cudaStream_t s) {
        if (!a) { cudaEventCreate(&a); cudaEventCreate(&b); }
        cudaEventRecord(a, s);
        armed = false;
    }
    void stop(cudaStream_t s) { cudaEventRecord(b, s); armed = true; }
    double ms() {
        if (!armed) return 0.0;
        float t = 0;
        if (cudaEventSynchronize(b) != cudaSuccess) return 0.0;
        cudaEventElapsedTime(&t, a, b);
        return t;
    }
    void destroy() { if (a) { cudaEventDestroy(a); cudaEventDestroy(b); a = b = nullptr; } }
};
#else
inline int radix_sort_pairs(Status &, Buf<uint8_t> &, const uint64_t *kin, uint64_t *kout,
                            const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit, cudaStream_t) {
    std::vector<uint32_t> o((size_t)n);
    std::iota(o.begin(), o.end(), 0u);
    uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
    std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (int64_t i = 0; i < n; i++) { kout[i] = kin[o[i]]; vout[i] = vin[o[i]]; }
    return 0;
}
inline int radix_sort_pairs32(Status &, Buf<uint8_t> &, const uint32_t *kin, uint32_t *kout,
                              const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit, cudaStream_t) {
    std::vector<uint32_t> o((size_t)n);
    std::iota(o.begin(), o.end(), 0u);
    uint32_t mask = end_bit >= 32 ? ~0u : ((1u << end_bit) - 1);
    std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (int64_t i = 0; i < n; i++) { kout[i] = kin[o[i]]; vout[i] = vin[o[i]]; }
    return 0;
}
inline int radix_sort_keys(Status &, Buf<uint8_t> &, const uint64_t *kin, uint64_t *kout, int64_t n, int end_bit, cudaStream_t) {
    uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
    std::vector<uint64_t> o(kin, kin + n);
    std::stable_sort(o.begin(), o.end(), [&](uint64_t a, uint64_t b) { return (a & mask) < (b & mask); });
    for (int64_t i = 0; i < n; i++) kout[i] = o[i];
    return 0;
}
inline int exclusive_sum_u32(Status &, Buf<uint8_t> &, uint32_t *in, uint32_t *out, int64_t n, cudaStream_t) {
    uint32_t acc = 0;
    for (int64_t i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = acc; acc += v; }
    out[n] = acc;
    return 0;
}
inline int exclusive_sum_u64(Status &, Buf<uint8_t> &, uint64_t *in, uint64_t *out, int64_t n, cudaStream_t) {
    uint64_t acc = 0;
    for (int64_t i = 0; i < n; i++) { uint64_t v = in[i]; out[i] = acc; acc += v; }
    out[n] = acc;
    return 0;
}
inline int inclusive_max_i32(Status &, Buf<uint8_t> &, int32_t *a, int64_t n, cudaStream_t) {
    for (int64_t i = 1; i < n; i++) a[i] = a[i] > a[i - 1] ? a[i] : a[i - 1];
    return 0;
}
struct Timer {
    std::chrono::steady_clock::time_point a, b;
    bool armed = false;
    void start(cudaStream_t) { a = std::chrono::steady_clock::now(); armed = false; }
    void stop(cudaStream_t) { b = std::chrono::steady_clock::now(); armed = true; }
    double ms() { return armed ? std::chrono::duration<double, std::milli>(b - a).count() : 0.0; }
    void destroy() {}
};
#endif

}  // namespace icl
