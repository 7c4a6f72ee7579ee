// Micro-benchmark: cost of warp min-reductions on sm_100a (developer tool, not part of the library).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t *out, int iters) {
    uint32_t v = threadIdx.x * 2654435761u + blockIdx.x, acc = 0xFFFFFFFFu;
    uint32_t w[8];
    for (int i = 0; i < 8; i++) w[i] = v * (i + 3);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {            // CREDUX chain-free: 8 independent reductions
                uint32_t r = __reduce_min_sync(0xffffffffu, w[u] ^ it);
                if ((threadIdx.x & 31) == u) acc = min(acc, r);
            } else if (MODE == 1) {     // butterfly shuffle min (5 shfl)
                uint32_t x = w[u] ^ it;
                for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
                acc = min(acc, x);
            } else if (MODE == 2) {     // single shfl + min
                uint32_t x = w[u] ^ it;
                x = min(x, __shfl_xor_sync(0xffffffffu, x, 16));
                acc = min(acc, x);
            } else if (MODE == 3) {     // dependent CREDUX chain (latency)
                acc = __reduce_min_sync(0xffffffffu, acc ^ w[u]);
            } else if (MODE == 4) {     // popc only
                acc += __popc(w[u] ^ it ^ acc);
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (uint32_t)(t1 - t0);
}
template <int MODE> void run(const char *name, int blocks, int threads) {
    uint32_t *d; cudaMalloc(&d, (blocks * threads + 1) * 4);
    const int iters = 2000;
    k<MODE><<<blocks, threads>>>(d, iters); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint32_t cyc; cudaMemcpy(&cyc, d + blocks * threads, 4, cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8;
    printf("%-28s blocks=%4d thr=%4d: %.1f cycles per op per warp (block0 clock), %.3f ms\n", name, blocks, threads, cyc / ops, ms);
    cudaFree(d);
}
int main() {
    for (int thr : {32, 128, 512}) {
        run<0>("credux independent", 148, thr);
        run<3>("credux dependent chain", 148, thr);
        run<1>("shfl butterfly (5 shfl)", 148, thr);
        run<2>("single shfl+min", 148, thr);
        run<4>("popc dependent", 148, thr);
    }
    return 0;
}
