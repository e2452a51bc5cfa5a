// Probe: bandwidth of SM-initiated stores to page-locked host memory (what the streamed export of the DP kernel does),
// by store width and by number of CTAs, against a copy-engine D2H of the same size.
#include <cuda_runtime.h>
#include <cstdio>
template <int W>
__global__ void k_write(int *dst, size_t n_int, int val) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * W, stride = (size_t)gridDim.x * blockDim.x * W;
    for (; i + W <= n_int; i += stride) {
        if (W == 1) dst[i] = val;
        if (W == 2) *(int2 *)(dst + i) = make_int2(val, val);
        if (W == 4) *(int4 *)(dst + i) = make_int4(val, val, val, val);
    }
}
int main() {
    const size_t bytes = 16u << 20, n = bytes / 4;
    int *h, *d; cudaHostAlloc(&h, bytes, cudaHostAllocDefault); cudaMalloc(&d, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float t;
    for (int ctas : {148, 722, 2960}) for (int thr : {32, 64, 256}) {
        float best[3] = {1e9f, 1e9f, 1e9f};
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(a); k_write<1><<<ctas, thr>>>(h, n, r); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); if (t < best[0]) best[0] = t;
            cudaEventRecord(a); k_write<2><<<ctas, thr>>>(h, n, r); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); if (t < best[1]) best[1] = t;
            cudaEventRecord(a); k_write<4><<<ctas, thr>>>(h, n, r); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); if (t < best[2]) best[2] = t;
        }
        printf("ctas %4d x %3d thr: 4B/lane %.1f GB/s  8B/lane %.1f GB/s  16B/lane %.1f GB/s\n", ctas, thr, bytes / best[0] / 1e6, bytes / best[1] / 1e6, bytes / best[2] / 1e6);
    }
    float bestc = 1e9f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(a); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); if (t < bestc) bestc = t; }
    printf("copy engine D2H: %.1f GB/s\n", bytes / bestc / 1e6);
    return 0;
}
