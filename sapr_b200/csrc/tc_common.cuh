// tc_common.cuh -- PTX wrappers and geometry shared by the tensor-core kernels (viterbi_tc.cu, estep_tc.cu):
// mbarrier / bulk-copy / tcgen05 (alloc, mma, ld, st, commit, fences), packed fp32x2 arithmetic, fp16 hi/lo packing,
// the no-swizzle K-major shared-memory matrix descriptor.
#pragma once
#include <cuda.h>          /* CUtensorMap type + enums only: the encoder is fetched through cudaGetDriverEntryPoint */
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

#define TC_ROWS 128
#define TC_GROUPS 4
#define TC_WORKERS (TC_ROWS * TC_GROUPS)
#define TC_WORKER_WARPS (TC_WORKERS / 32)
#define TC_LOADERS 3
#define TC_THREADS (TC_WORKERS + 32 + 32 * TC_LOADERS)    /* + MMA warp + bulk-copy warps */
#define TC_MAX_STAGES 4
#define TC_TRQ 8          /* float4 per model in the transition block (k_prepare_tc) */

// ------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// bounded wait: a protocol bug traps (error to the host) instead of hanging the GPU.  try_wait suspends the
// thread in hardware up to the time hint, so a waiting warp costs almost no issue slots.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef SAPR_WAIT_HINT
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.gt.u32 p, n, 4000000;\n\t"
        "@p trap;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}"
        ::"r"(bar), "r"(parity), "r"(2000u) : "memory");
#else
    // no suspend-time hint: a failed try is TRYWAIT + branch (2 SASS instructions); with the hint the compiler adds a
    // NANOSLEEP.SYNCS and a re-check per try, and every barrier event in the CTA wakes the sleeper (measured: 23 % of the
    // Viterbi kernel's issued instructions were such polls)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.gt.u32 p, n, 16000000;\n\t"
        "@p trap;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
#endif
}
// non-blocking phase test, and one suspending try (returns after the hardware's time limit at the latest)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(e));
    return e != 0;
}
// keeps a per-thread constant in its register (stops the compiler from re-deriving it from %tid in the hot loop)
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// TMA tensor-map tile load (3-D box) global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// TMA tile prefetch into L2 only: extends the bytes in flight beyond what the shared-memory ring holds
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap *tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// bulk prefetch global -> L2: extends the bytes in flight beyond what the shared-memory ring can hold
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::f16, fp32 accumulate
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two lanes per instruction)
__device__ __forceinline__ unsigned long long f2_as_u64(float2 a) { return *reinterpret_cast<unsigned long long *>(&a); }
__device__ __forceinline__ float2 u64_as_f2(unsigned long long a) { return *reinterpret_cast<float2 *>(&a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)), "l"(f2_as_u64(c)));
    return u64_as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(d);
}
// (lo16, hi16) = (fp16(a.x), fp16(a.y)), saturating to the largest finite half instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_h2(float2 a) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a.y), "f"(a.x));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t h) { return __half22float2(*reinterpret_cast<__half2 *>(&h)); }
// a - (fp32)h, per element, as ONE mixed-precision FMA each (FHFMA: fp32 = half * half + fp32 with the constant -1):
// the residual of an fp16 split without converting the hi part back to fp32 first
__device__ __forceinline__ float2 residual_h2(float2 a, uint32_t h) {
    float2 r;
    asm("{\n\t.reg .b16 h0, h1, m1;\n\tmov.b32 {h0, h1}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, h0, m1, %3;\n\tfma.rn.f32.f16 %1, h1, m1, %4;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "r"(h), "f"(a.x), "f"(a.y));
    return r;
}

// K-major, no-swizzle ("interleave") shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// core matrix = 8 rows x 16 bytes stored contiguously (128 B); LBO = byte distance between the two core
// matrices of one K = 16 step, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (bits 61-63 = 0)
}

// ------------------------------------------------------------------------------------------------
// host: 3-D fp32 tensor map (innermost dimension first).  The encoder lives in libcuda; it is looked up at run time so
// the library links against cudart only.  Boxes may reach past the tensor: out-of-range elements arrive as zeros.
typedef CUresult (*sapr_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline int sapr_tmap_f32_3d(sapr_ctx *ctx, CUtensorMap *out, const void *base, const uint64_t (&dim)[3],
                                   const uint64_t (&stride_bytes)[2], const uint32_t (&box)[3]) {
    static sapr_tmap_encode_fn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        SAPR_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) SAPR_FAIL(ctx, SAPR_E_CUDA, "cuTensorMapEncodeTiled not available");
        enc = (sapr_tmap_encode_fn)fn;
    }
    const cuuint64_t gd[3] = {dim[0], dim[1], dim[2]};
    const cuuint64_t gs[2] = {stride_bytes[0], stride_bytes[1]};
    const cuuint32_t bx[3] = {box[0], box[1], box[2]};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SAPR_FAIL(ctx, SAPR_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return SAPR_OK;
}
