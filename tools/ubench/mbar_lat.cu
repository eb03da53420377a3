// Micro-benchmark (tuning aid): latency of a phase check on an mbarrier whose phase has completed long ago:
// mbarrier.try_wait.parity vs mbarrier.test_wait.parity, one warp, dependent chain (the predicate feeds the next address).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_lat mbar_lat.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out) {
    __shared__ uint64_t bar[2];
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b + 8));
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b));
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b + 8));
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        for (int mode = 0; mode < 3; mode++) {
            uint32_t a = b, acc = 0;
            long long t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < 256; i++) {
                uint32_t ok;
                if (mode == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a) : "memory");
                else if (mode == 1)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a) : "memory");
                else
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ok) : "r"(a) : "memory");
                acc += ok;
                a = b + ((ok & 1u) ^ 1u) * 8u + ((mode == 2) ? 0u : 0u);      // dependent address
            }
            long long t1 = clock64();
            if (threadIdx.x == 0) { out[mode] = t1 - t0; out[4 + mode] = acc; }
        }
    }
}
int main() {
    long long *d, h[8];
    cudaMalloc(&d, 64);
    k<<<1, 64>>>(d); cudaDeviceSynchronize();
    k<<<1, 64>>>(d); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("try_wait %.1f cycles, test_wait %.1f cycles, ld.shared %.1f cycles per dependent check (ok counts %lld %lld)\n", h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[4], h[5]);
    return 0;
}
