// Micro-benchmark (tuning aid): throughput of tcgen05.ld / tcgen05.st (32x32b shapes) per SM, by instruction width and
// number of issuing warps; several accesses in flight per warp (one wait per batch of 8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, long long *out, uint32_t *sink) {
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tm + (((uint32_t)(warp & 3) * 32) << 16) + 64u * (warp >> 2);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (MODE == 0) {
                    uint32_t v[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tb + 8u * (j & 3)) : "memory");
                    acc += v[0];
                } else if (MODE == 1) {
                    uint32_t v[32];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(tb + 32u * (j & 1)) : "memory");
                    acc += v[0] + v[31];
                } else if (MODE == 2) {
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb + 4u * j), "r"(acc), "r"(acc), "r"(acc), "r"(acc) : "memory");
                } else if (MODE == 3) {
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                                 ::"r"(tb + 16u * (j & 3)), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc), "r"(acc),
                                 "r"(acc), "r"(acc), "r"(acc), "r"(acc) : "memory");
                }
            }
            if (MODE < 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    long long t1 = clock64();
    if (acc == 0x1234567u) sink[0] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int MODE> void run(const char *name, int bytes, long long *d, uint32_t *sink) {
    for (int nw : {1, 4, 8, 16}) {
        const int iters = 200;
        long long h = 0;
        for (int rep = 0; rep < 2; rep++) { k<MODE><<<1, 512>>>(nw, iters, d, sink); cudaDeviceSynchronize(); }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double ops = (double)nw * iters * 8;
        printf("%-12s %2d warps: %6.1f cycles per warp-instruction (SM-wide %5.2f), %6.1f B/cycle\n", name, nw, (double)h / (iters * 8), (double)h / ops, ops * bytes / (double)h);
    }
}
int main() {
    long long *d; uint32_t *sink;
    cudaMalloc(&d, 64); cudaMalloc(&sink, 16);
    run<0>("LDTM.x8", 1024, d, sink); run<1>("LDTM.x32", 4096, d, sink); run<2>("STTM.x4", 512, d, sink); run<3>("STTM.x16", 2048, d, sink);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return 0;
}
