// Micro-benchmark: what does a random 32-byte probe cost in DRAM bytes on this part?
// Each thread reads N random 32-byte-aligned records from a buffer much larger than L2.
// Variants select the load instruction.  Run under ncu for dram__bytes_read.sum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int V>
__global__ void gather(const uint4 *buf, size_t nrec, int iters, unsigned *sink) {
    unsigned x = blockIdx.x * blockDim.x + threadIdx.x + 12345u;
    unsigned acc = 0;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        size_t r = ((size_t)x * 2654435761u >> 7) % nrec;
        const uint4 *p = buf + 2 * r;
        unsigned a0, a1, a2, a3, b0, b1, b2, b3;
        if (V == 0) asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
        if (V == 1) asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
        if (V == 2) { asm volatile("ld.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p));
                      asm volatile("ld.global.v4.b32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p)); }
        if (V == 3) { asm volatile("ld.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p)); b0 = b1 = b2 = b3 = 0; }  // 16 B only
        if (V == 4) asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
        if (V == 5) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
        acc += a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
    }
    if (acc == 0x12345678u) *sink = acc;
}
int main(int argc, char **argv) {
    size_t bytes = (size_t)4 << 30;
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) printf("set L2 fetch granularity %d: %d\n", gran, (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity %zu\n", g);
    uint4 *buf; unsigned *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4); cudaMemset(buf, 1, bytes);
    size_t nrec = bytes / 32;
    const int grid = 148 * 8, block = 128, iters = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
#define RUN(V) cudaEventRecord(e0); gather<V><<<grid, block>>>(buf, nrec, iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); \
    cudaEventElapsedTime(&ms, e0, e1); printf("variant %d: %.3f ms, %.1f M probes, %.1f GB useful/s\n", V, ms, grid * block * (double)iters / 1e6, grid * block * (double)iters * 32 / ms / 1e6);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("probes per launch %d, useful bytes %.1f MB\n", grid * block * iters, grid * block * (double)iters * 32 / 1e6);
    return 0;
}
