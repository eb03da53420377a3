// Correctness probe (tuning aid): one fp16 shared-memory image [16 row groups][10 K chunks][8 rows][8 halves] read by
// tcgen05.mma two ways -- as a K-major A operand (rows = M, emission product) and as an MN-major A operand (features = M,
// rows = K: the statistics product Gamma^T . X of the fused E-step) -- plus an MN-major B operand, against the CPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mn_major_test mn_major_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
#define NCK 10
#define RG_STRIDE (NCK * 128)
// img: A image 16*NCK*128 B; w: [2][NCK][8][8] halves (K-major B, 16 rows n); gam: [16][2][8][8] halves (MN-major B)
__global__ void k(const __half *img, const __half *w, const __half *gam, float *d1, float *d2) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tm;
    __half *sA = reinterpret_cast<__half *>(smem);
    __half *sW = reinterpret_cast<__half *>(smem + 16 * RG_STRIDE);
    __half *sG = reinterpret_cast<__half *>(smem + 16 * RG_STRIDE + 2 * NCK * 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * RG_STRIDE / 2; i += blockDim.x) sA[i] = img[i];
    for (int i = threadIdx.x; i < 2 * NCK * 64; i += blockDim.x) sW[i] = w[i];
    for (int i = threadIdx.x; i < 16 * 2 * 64; i += blockDim.x) sG[i] = gam[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tm;
    if (warp == 0) {
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        if (elected) {
            const uint32_t idesc_k = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_mn = idesc_k | (1u << 15) | (1u << 16);
            // product 1: D1[r][n] = sum_k A[r][k] W[n][k]; A K-major (LBO = 128 between the K chunks, SBO = row-group stride)
            for (int ks = 0; ks < NCK / 2; ks++)
                mma_ss(tb, make_desc(smem_u32(sA) + ks * 256, 128, RG_STRIDE), make_desc(smem_u32(sW) + ks * 256, 128, NCK * 128), idesc_k, ks > 0);
            // product 2: D2[f][n] = sum_r A[r][f] G[r][n]; A MN-major (LBO = K-group (row-group) stride, SBO = MN-group (chunk) stride),
            // B MN-major: [row group][n group][8 rows][8 n]
            for (int ks = 0; ks < 8; ks++)
                mma_ss(tb + 16, make_desc(smem_u32(sA) + 2 * ks * RG_STRIDE, RG_STRIDE, 128), make_desc(smem_u32(sG) + 2 * ks * 256, 256, 128), idesc_mn, ks > 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
    }
    {
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[32];
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(ta) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int r = warp * 32 + lane;
    for (int n = 0; n < 16; n++) { d1[r * 16 + n] = __uint_as_float(v[n]); d2[r * 16 + n] = __uint_as_float(v[16 + n]); }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb));
}

int main() {
    const int K = 8 * NCK;
    std::vector<float> A(128 * K), W(16 * K), G(128 * 16);
    for (int r = 0; r < 128; r++) for (int kk = 0; kk < K; kk++) A[r * K + kk] = (float)((r * 7 + kk * 3) % 11 - 5) / 8.f;
    for (int n = 0; n < 16; n++) for (int kk = 0; kk < K; kk++) W[n * K + kk] = (float)((n * 5 + kk) % 7 - 3) / 4.f;
    for (int r = 0; r < 128; r++) for (int n = 0; n < 16; n++) G[r * 16 + n] = (float)((r * 3 + n * 2) % 9) / 16.f;
    std::vector<__half> img(16 * NCK * 64), w(2 * NCK * 64), gam(16 * 2 * 64);
    for (int r = 0; r < 128; r++) for (int kk = 0; kk < K; kk++) img[((r / 8) * NCK + kk / 8) * 64 + (r % 8) * 8 + kk % 8] = __float2half(A[r * K + kk]);
    for (int n = 0; n < 16; n++) for (int kk = 0; kk < K; kk++) w[((n / 8) * NCK + kk / 8) * 64 + (n % 8) * 8 + kk % 8] = __float2half(W[n * K + kk]);
    for (int r = 0; r < 128; r++) for (int n = 0; n < 16; n++) gam[((r / 8) * 2 + n / 8) * 64 + (r % 8) * 8 + n % 8] = __float2half(G[r * 16 + n]);
    __half *di, *dw, *dg; float *d1, *d2;
    cudaMalloc(&di, img.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dg, gam.size() * 2);
    cudaMalloc(&d1, 128 * 16 * 4); cudaMalloc(&d2, 128 * 16 * 4);
    cudaMemcpy(di, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dg, gam.data(), gam.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 16 * RG_STRIDE + 2 * NCK * 128 + 16 * 256 + 4096;   // + slack the M = 128 MN-major read runs into
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 128, smem>>>(di, dw, dg, d1, d2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> h1(128 * 16), h2(128 * 16);
    cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int r = 0; r < 128; r++) for (int n = 0; n < 16; n++) {
        double s = 0; for (int kk = 0; kk < K; kk++) s += (double)A[r * K + kk] * W[n * K + kk];
        e1 = fmax(e1, fabs(s - h1[r * 16 + n]));
    }
    for (int f = 0; f < K; f++) for (int n = 0; n < 16; n++) {
        double s = 0; for (int r = 0; r < 128; r++) s += (double)A[r * K + f] * G[r * 16 + n];
        e2 = fmax(e2, fabs(s - h2[f * 16 + n]));
    }
    printf("mn_major_test: K-major product max err %.3g, MN-major product max err %.3g (%s)\n", e1, e2, (e1 < 1e-3 && e2 < 1e-3) ? "OK" : "MISMATCH");
    if (e2 >= 1e-3) { for (int f = 0; f < 4; f++) { for (int n = 0; n < 16; n++) printf("%7.3f ", h2[f * 16 + n]); printf("\n"); } }
    return 0;
}
