// Micro-benchmark (tuning aid): cycles per tcgen05.mma (kind::f16, M = 128, N = 96, K = 16, A from TMEM, B from shared memory)
// while other warps of the CTA keep the TMEM load / store ports or the shared-memory pipe busy.
//   bg 0: nothing   1: 16 warps tcgen05.ld.x8   2: 8 warps tcgen05.st.x4   3: both   4: 8 warps LDS.128 streaming   5: 16 warps FMA/ALU loop
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contend umma_contend.cu
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

__global__ void __launch_bounds__(26 * 32, 1) k(int N, int bg, int iters, long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tm;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 24) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tm;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    if (warp == 24) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sbo = 10 * 128u;
        const uint64_t bdesc = make_desc(smem_u32(smem), 128, sbo);
        long long t0 = 0, t1 = 0, t2 = 0;
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        if (elected) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 5) {
#pragma unroll
                for (int ks = 0; ks < 5; ks++) mma_ts(tb + (uint32_t)(((i / 15) & 1) * (N <= 96 ? 96 : 0)), tb + 400 + 16 * ks, bdesc + 16 * ks, idesc, 1);
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)));
            t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp < 16 && (bg == 1 || bg == 3)) {
        uint32_t acc = 0;
        long long n = 0;
        while (!stop) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tb + lane_sel + 8u * (warp >> 2)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[7]; n++;
        }
        if (acc == 0x12345u) sink[0] = (float)acc;
        if (lane == 0) out[2 + warp] = n;
    } else if (warp >= 16 && warp < 24 && (bg == 2 || bg == 3)) {
        long long n = 0;
        while (!stop) {
#pragma unroll
            for (int c = 0; c < 5; c++)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb + lane_sel + 272u + 16u * c + 40u * ((warp - 16) >> 2)),
                             "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            n += 5;
        }
        if (lane == 0) out[2 + warp] = n;
    } else if (warp >= 16 && warp < 24 && bg == 4) {
        float acc = 0.f;
        long long n = 0;
        uint32_t a = smem_u32(smem) + 32768 + (warp - 16) * 4096 + lane * 176 % 4096;
        while (!stop) {
#pragma unroll
            for (int c = 0; c < 5; c++) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a + 16u * c));
                acc += v.x + v.w;
            }
            n += 5;
        }
        if (acc == 123.f) sink[0] = acc;
        if (lane == 0) out[2 + warp] = n;
    } else if (warp < 16 && bg == 5) {
        float x = (float)lane, y = 1.0f;
        uint32_t s = lane;
        long long n = 0;
        while (!stop) {
#pragma unroll
            for (int c = 0; c < 16; c++) { x = fmaf(x, 1.0001f, y); s = __funnelshift_l(__float_as_uint(x), s, 1); y = fmaxf(y, x * 0.5f); }
            n += 48;
        }
        if (x == 123.f && s == 7) sink[0] = x + y;
        if (lane == 0) out[2 + warp] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 24) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    long long *d, h[32];
    float *sink;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&sink, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 3000;
    for (int bg = 0; bg < 6; bg++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(d, 0, sizeof(h));
            k<<<1, 26 * 32, 100 * 1024>>>(96, bg, iters, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("bg %d: %s\n", bg, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long nb = 0;
        for (int i = 2; i < 26; i++) nb += h[i];
        printf("bg %d: issue %.1f cyc/mma, complete %.1f cyc/mma; background warp-ops %lld (%.2f per cycle)\n", bg, (double)h[0] / iters,
               (double)h[1] / iters, nb, (double)nb / (double)h[1]);
    }
    // cost of one product as a function of N (no background work; accumulator columns 0 .. N - 1, A at column 256)
    const int ns[] = {16, 32, 48, 64, 96, 128, 192};
    for (int N : ns) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(d, 0, sizeof(h));
            k<<<1, 26 * 32, 100 * 1024>>>(N, 0, iters, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N %3d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", N, (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}
