// Micro-benchmark (tuning aid): cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread,
// A from TMEM (TS) or shared memory (SS), B from shared memory (K-major, no swizzle), for several N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

// mode 0: TS same accumulator; 1: SS same accumulator; 2: TS alternating two accumulators; 3: TS, B descriptor fixed (same B tile)
__global__ void k(int N, int mode, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tm;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int nck = 10;
        const uint32_t sbo = nck * 128u;
        const uint64_t bdesc = make_desc(smem_u32(smem), 128, sbo);
        const uint64_t adesc = make_desc(smem_u32(smem) + 16384, 128, sbo);
        long long t0 = 0, t1 = 0, t2 = 0;
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        if (elected) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 5) {
#pragma unroll
                for (int ks = 0; ks < 5; ks++) {
                    const uint32_t d = tb + ((mode == 2) ? (uint32_t)((ks & 1) * 128) : 0u);
                    if (mode == 1) mma_ss(d, adesc + 16 * ks, bdesc + 16 * ks, idesc, 1);
                    else mma_ts(d, tb + 256 + 16 * ks, bdesc + (mode == 3 ? 0 : 16 * ks), idesc, 1);
                }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)));
            t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 600;
    for (int mode = 0; mode < 4; mode++)
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            for (int rep = 0; rep < 2; rep++) {
                k<<<1, 128, 64 * 1024>>>(N, mode, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode %d (%s) N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", mode,
                   mode == 0 ? "TS" : mode == 1 ? "SS" : mode == 2 ? "TS 2 acc" : "TS same B", N, (double)h[0] / iters,
                   (double)h[1] / iters, 128 * N / 256);
        }
    return 0;
}
