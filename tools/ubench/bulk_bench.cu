// Micro-benchmark (tuning aid): issue cost of per-lane cp.async.bulk (global -> shared) as a function of the copy size and the
// number of copy warps; every lane copies its own row (like the feature ring of the tensor-core kernels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bench bulk_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each of nw warps issues `rounds` warp-instructions of 32 per-lane bulk copies of `bytes`; one mbarrier per warp per round pair
__global__ void k(const float *src, size_t row_stride_f, int bytes, int rounds, int nw, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[threadIdx.x])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
    __syncthreads();
    if (warp < nw) {
        const uint32_t b = smem_u32(&bar[warp]);
        const uint32_t dst = smem_u32(smem) + (uint32_t)(warp * 32 + lane) * (uint32_t)(bytes + 16);
        const float *s = src + (size_t)(blockIdx.x * 1024 + warp * 32 + lane) * row_stride_f;
        long long t0 = clock64(), tiss = 0;
        uint32_t ph = 0;
        for (int r = 0; r < rounds; r++) {
            long long a = clock64();
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"((uint32_t)bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(s + (size_t)r * (bytes / 4)), "r"((uint32_t)bytes), "r"(b) : "memory");
            tiss += clock64() - a;
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(b), "r"(ph));
            ph ^= 1u;
        }
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[2 * warp] = tiss; out[2 * warp + 1] = t1 - t0; }
    }
}

// issue only (no wait between rounds; depth copies in flight per lane, waits every `depth` rounds)
__global__ void k2(const float *src, size_t row_stride_f, int bytes, int rounds, int nw, int depth, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[threadIdx.x])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    __syncthreads();
    if (warp < nw) {
        const uint32_t b = smem_u32(&bar[warp]);
        const uint32_t dst = smem_u32(smem) + (uint32_t)(warp * 32 + lane) * (uint32_t)(bytes + 16);
        const float *s = src + (size_t)(blockIdx.x * 1024 + warp * 32 + lane) * row_stride_f;
        long long t0 = clock64();
        uint32_t ph = 0;
        for (int r = 0; r < rounds; r += depth) {
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"((uint32_t)(bytes * depth)) : "memory");
            for (int d = 0; d < depth; d++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(s + (size_t)(r + d) * (bytes / 4)), "r"((uint32_t)bytes), "r"(b) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(b), "r"(ph));
            ph ^= 1u;
        }
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    }
}

int main() {
    const size_t row_f = 200 * 40;                 // one utterance: 200 frames x 40 floats
    const size_t rows = 148 * 1024;
    float *src; long long *d, h[64];
    cudaMalloc(&src, rows * row_f * 4);
    cudaMemset(src, 0, rows * row_f * 4);
    cudaMalloc(&d, 64 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int bytes : {160, 320, 640, 1280})
        for (int nw : {1, 3, 6, 8}) {
            if ((size_t)nw * 32 * (bytes + 16) > 200 * 1024) continue;
            const int rounds = 6400 / bytes * 5;       // <= 200 frames
            for (int rep = 0; rep < 2; rep++) k<<<148, 1024, 200 * 1024>>>(src, row_f, bytes, rounds, nw, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
            printf("bytes %4d warps %d: issue (arrive+copy) %.0f cyc per warp-instruction, round trip %.0f cyc  [warp 0]\n", bytes, nw,
                   (double)h[0] / rounds, (double)h[1] / rounds);
        }
    for (int bytes : {160, 320, 640})
        for (int nw : {1, 3, 6})
            for (int depth : {2, 4}) {
                const int rounds = 6400 / bytes * 5 / depth * depth;
                if ((size_t)nw * 32 * (bytes + 16) > 200 * 1024) continue;
                for (int rep = 0; rep < 2; rep++) k2<<<148, 1024, 200 * 1024>>>(src, row_f, bytes, rounds, nw, depth, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
                printf("bytes %4d warps %d depth %d (same dst, throughput only): %.0f cyc per warp-instruction of 32 copies\n", bytes, nw, depth,
                       (double)h[0] / rounds);
            }
    return 0;
}
