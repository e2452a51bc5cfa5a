// PCIe probe: D2H / H2D of the sizes the hot path moves, pinned memory, non-blocking stream, CUDA events.
#include <cuda_runtime.h>
#include <cstdio>
int main() {
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const size_t sizes[] = {1u << 20, 4u << 20, 8u << 20, 15u << 20, 32u << 20};
    void *d, *h, *h2; cudaMalloc(&d, 64u << 20); cudaHostAlloc(&h, 64u << 20, cudaHostAllocDefault);
    cudaMallocHost(&h2, 64u << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (size_t sz : sizes) {
        float best_d2h = 1e9, best_h2d = 1e9, t;
        for (int r = 0; r < 10; r++) {
            cudaEventRecord(a, s); cudaMemcpyAsync(h, d, sz, cudaMemcpyDeviceToHost, s); cudaEventRecord(b, s); cudaEventSynchronize(b);
            cudaEventElapsedTime(&t, a, b); if (t < best_d2h) best_d2h = t;
            cudaEventRecord(a, s); cudaMemcpyAsync(d, h, sz, cudaMemcpyHostToDevice, s); cudaEventRecord(b, s); cudaEventSynchronize(b);
            cudaEventElapsedTime(&t, a, b); if (t < best_h2d) best_h2d = t;
        }
        printf("%5.1f MiB  d2h %.3f ms (%.1f GB/s)  h2d %.3f ms (%.1f GB/s)\n", sz / 1048576.0, best_d2h, sz / best_d2h / 1e6, best_h2d, sz / best_h2d / 1e6);
    }
    return 0;
}
