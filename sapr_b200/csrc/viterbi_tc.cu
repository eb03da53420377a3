// viterbi_tc.cu -- batched Viterbi with the diagonal-Gaussian emission on the 5th-gen tensor cores.
//
// Same contract as k_viterbi_fused (custom_hmm.py:462-514 x all models + decoder.py:42-47), different
// machine mapping.  The emission of one frame against all M*8 states is the dense contraction
//      E[r, n] = sum_k A[r, k] * W[n, k],   A[r, :] = [x', x'^2, 1] of utterance r at frame t,
// where x' = (x - g) * s is the feature standardised by a per-dimension centre/scale derived from the
// model set (keeps the expanded quadratic well conditioned) and W[n, :] = [2 mu' h', -h', c - sum mu'^2 h'].
// One MMA tile = 128 utterances AT THE SAME FRAME INDEX: tcgen05.mma writes row r to TMEM lane r, and
// tcgen05.ld.32x32b hands thread r exactly the emissions of ITS utterance -- no transposition between the
// tensor-core tile and the register-resident left-to-right recursion.
//
// Precision: operands are fp16 hi/lo splits (x' = hi + lo exactly to 22 bits, likewise W); the three products
// hi*Whi + lo*Whi + hi*Wlo are one K = 3*Kh accumulation chain in fp32 TMEM (Kh = 80 for D = 39): emission
// error ~1e-5 absolute, the same class as the fp32 SIMT kernel.
//
// CTA = 9 warps: warps 0-7 are workers (thread = (row r, column group g): converts half of row r's next
// frame into the A tile, then runs the recursion for its half of the models), warp 8 issues the MMAs.
// Pipeline per frame f (stage = f & 1):  workers write A[f+1] -> mbarrier A_full -> MMA warp issues
// 3*nck/2 tcgen05.mma into TMEM buffer (f+1)&1 -> tcgen05.commit -> mbarrier acc_full -> workers tcgen05.ld,
// recursion, mbarrier acc_empty.  The MMAs of frame f+1 overlap the recursion of frame f.
#include <cuda_fp16.h>

#include "common.cuh"

#define TC_WORKERS 256
#define TC_THREADS 288
#define TC_ROWS 128

// ------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps (error to the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
// model-set preparation: standardisation (g, s) and the fp16 hi/lo weight image in its shared-memory layout
//   wimg[2][ncols/8][nck][8][8] halves  (hi plane, lo plane); n = m*8 + state, k = 2*dim (x') / 2*dim+1 (x'^2)
//   gs[4*nck] float2 (g, s) for real dims;  trp[M][5] float4 transition block
__global__ void k_prepare_tc(int M, int S, int D, int nck, int ncols, const double *__restrict__ mean,
                             const double *__restrict__ var, const double *__restrict__ la, const double *__restrict__ lb,
                             __half *__restrict__ wimg, float2 *__restrict__ gs, float4 *__restrict__ trp) {
    extern __shared__ double s_gs[];   // [2][4*nck]
    const int N = S - 2, nd = 4 * nck;
    for (int d = threadIdx.x; d < nd; d += blockDim.x) {
        double g = 0.0, sc = 1.0;
        if (d < D) {
            double sm = 0.0;
            for (int m = 0; m < M; m++) for (int j = 1; j <= N; j++) sm += mean[((size_t)m * S + j) * D + d];
            g = sm / (M * N);
            double v = 0.0;
            for (int m = 0; m < M; m++)
                for (int j = 1; j <= N; j++) {
                    const double df = mean[((size_t)m * S + j) * D + d] - g;
                    v += var[((size_t)m * S + j) * D + d] + df * df;
                }
            v /= (M * N);
            sc = (v > 0 && v < 1e300) ? 4.0 / sqrt(v) : 1.0;
            g = (double)(float)g; sc = (double)(float)sc;   // the kernel standardises in fp32 with exactly these values
        }
        s_gs[d] = g; s_gs[nd + d] = sc;
        gs[d] = make_float2((float)g, (float)sc);
    }
    __syncthreads();
    const size_t plane = (size_t)(ncols / 8) * nck * 64;
    for (int idx = threadIdx.x; idx < ncols * nd; idx += blockDim.x) {
        const int n = idx / nd, d = idx % nd;
        const int m = n / 8, j = (n % 8) + 1;
        double wx = 0.0, wx2 = 0.0;
        if (m < M && j <= N) {
            const double *mu = mean + ((size_t)m * S + j) * D, *vr = var + ((size_t)m * S + j) * D;
            if (d < D) {
                const double g = s_gs[d], sc = s_gs[nd + d];
                const double hp = 0.5 / vr[d] / (sc * sc), mp = (mu[d] - g) * sc;
                wx = 2.0 * mp * hp; wx2 = -hp;
            } else if (d == D) {   // constant slot: A carries x' = 1 here
                double ld = 0.0, c2 = 0.0;
                for (int q = 0; q < D; q++) {
                    const double g = s_gs[q], sc = s_gs[nd + q];
                    const double hp = 0.5 / vr[q] / (sc * sc), mp = (mu[q] - g) * sc;
                    ld += log(vr[q]); c2 += mp * mp * hp;
                }
                wx = -0.5 * (D * SAPR_LOG2PI + ld) - c2; wx2 = 0.0;
            }
        }
        // element (n, k) of the image: group n/8, chunk k/8, row n%8, elem k%8; k = 2d, 2d+1
        const size_t o = ((size_t)(n / 8) * nck + d / 4) * 64 + (size_t)(n % 8) * 8 + (d % 4) * 2;
        const __half hx = __double2half(wx), hx2 = __double2half(wx2);
        wimg[o] = hx; wimg[o + 1] = hx2;
        wimg[plane + o] = __double2half(wx - (double)__half2float(hx));
        wimg[plane + o + 1] = __double2half(wx2 - (double)__half2float(hx2));
    }
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const double *a = la + (size_t)m * S, *b = lb + (size_t)m * S;
        // (stay_j, adv_j) pairs for states 1..8: adv_j = ln A[j-1, j] = lb[j-1], stay_j = ln A[j, j] = la[j]
        float v[20];
        for (int j = 1; j <= 8; j++) { v[2 * (j - 1)] = (j <= N) ? (float)a[j] : -INFINITY; v[2 * (j - 1) + 1] = (j <= N) ? (float)b[j - 1] : -INFINITY; }
        v[16] = (float)a[S - 1]; v[17] = (float)b[N]; v[18] = (float)b[0]; v[19] = 0.f;
        for (int q = 0; q < 5; q++) trp[(size_t)m * 5 + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
struct TcParams {
    const float *X; int ldx; const int64_t *offsets; int u0, nu, M, D, nck, ncols, first_frames;
    const __half *wimg; const float2 *gs; const float4 *trp;
    uint16_t *bp; int64_t Bpad; int maxT; double *scores; float *dbgE;   // dbgE: [sum_T][ncols] or null
};

template <int MG>
__global__ void __launch_bounds__(TC_THREADS, 1) k_viterbi_tc(const TcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int nck = p.nck, ncols = p.ncols, M = p.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- shared memory carve-up ----
    const uint32_t w_plane = (uint32_t)(ncols / 8) * nck * 128;            // bytes per W plane
    const uint32_t a_stage = 16u * 2u * nck * 128u;                         // 128 rows x (2*nck chunks) x 16 B
    unsigned char *sW = smem;                                               // hi plane, lo plane
    unsigned char *sA = sW + 2 * w_plane;                                   // 2 stages
    float4 *sTr = reinterpret_cast<float4 *>(sA + 2 * a_stage);             // [M][5]
    float2 *sGs = reinterpret_cast<float2 *>(sTr + (size_t)M * 5);          // [4*nck]
    uint64_t *sBar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(sGs + 4 * nck) + 15) & ~(uintptr_t)15);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + 8);
    const uint32_t barA_full = smem_u32(sBar + 0), barAcc_full = smem_u32(sBar + 2), barAcc_empty = smem_u32(sBar + 4);

    {   // stage the constant images
        const uint4 *src = reinterpret_cast<const uint4 *>(p.wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        for (uint32_t i = tid; i < 2 * w_plane / 16; i += TC_THREADS) dst[i] = src[i];
        for (int i = tid; i < M * 5; i += TC_THREADS) sTr[i] = p.trp[i];
        for (int i = tid; i < 4 * nck; i += TC_THREADS) sGs[i] = p.gs[i];
    }
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(barA_full + 8 * s, TC_WORKERS);
            mbar_init(barAcc_full + 8 * s, 1);
            mbar_init(barAcc_empty + 8 * s, TC_WORKERS);
        }
        fence_barrier_init();
    }
    // TMEM: two accumulator buffers of ncols columns
    uint32_t tcols = 32;
    while (tcols < 2u * ncols) tcols <<= 1;
    if (warp == 8) tmem_alloc(smem_u32(sTmem), tcols);
    fence_proxy_async();            // W image written with generic stores, read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;

    const int ntiles = (p.nu + TC_ROWS - 1) / TC_ROWS;
    uint32_t f = 0;                 // running frame counter of this CTA (pipeline stage / phase bookkeeping)
    __shared__ int s_Tt[2];         // per-tile frame count, ping-pong by tile iteration

    if (warp == 8) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        const uint32_t sW_hi = smem_u32(sW), sW_lo = sW_hi + w_plane, sA0 = smem_u32(sA);
        const uint32_t sboA = 2u * nck * 128u, sboW = (uint32_t)nck * 128u;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // frames this tile walks = max over its rows (computed identically by the workers)
            int Tt = 0;
            for (int r = lane; r < TC_ROWS; r += 32) {
                const int ul = tile * TC_ROWS + r;
                if (ul < p.nu) {
                    const int T = (int)(p.offsets[p.u0 + ul + 1] - p.offsets[p.u0 + ul]);
                    Tt = max(Tt, (p.first_frames > 0 && p.first_frames < T) ? p.first_frames : T);
                }
            }
            for (int o = 16; o > 0; o >>= 1) Tt = max(Tt, __shfl_xor_sync(0xffffffffu, Tt, o));
            for (int t = 0; t < Tt; t++, f++) {
                const uint32_t s = f & 1, ph = (f >> 1) & 1;
                mbar_wait(barA_full + 8 * s, ph);
                mbar_wait(barAcc_empty + 8 * s, ph ^ 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d_tmem = tmem_base + s * (uint32_t)ncols;
                    const uint32_t aBase = sA0 + s * a_stage;
                    uint32_t acc = 0;
                    for (int ks = 0; ks < nck / 2; ks++) {      // (hi + lo) * W_hi ... hi block
                        umma_f16(d_tmem, make_desc(aBase + ks * 256, 128, sboA), make_desc(sW_hi + ks * 256, 128, sboW), idesc, acc);
                        acc = 1;
                    }
                    for (int ks = 0; ks < nck / 2; ks++)        // lo block of A against W_hi
                        umma_f16(d_tmem, make_desc(aBase + (nck + 2 * ks) * 128, 128, sboA), make_desc(sW_hi + ks * 256, 128, sboW), idesc, 1);
                    for (int ks = 0; ks < nck / 2; ks++)        // hi block of A against W_lo
                        umma_f16(d_tmem, make_desc(aBase + ks * 256, 128, sboA), make_desc(sW_lo + ks * 256, 128, sboW), idesc, 1);
                    umma_commit(barAcc_full + 8 * s);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== workers =====================
        const int g = warp >> 2;                     // column group (which half of the models)
        const int r = (warp & 3) * 32 + lane;        // row of the tile = TMEM lane
        const int mg0 = (M + 1) / 2;
        const int mbeg = g == 0 ? 0 : mg0, mcnt = g == 0 ? mg0 : M - mg0;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int cbeg = g == 0 ? 0 : (nck + 1) / 2, cend = g == 0 ? (nck + 1) / 2 : nck;   // chunks this thread converts
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int ul = tile * TC_ROWS + r;
            const bool live = ul < p.nu;
            int64_t off = 0;
            int Te = 0;
            if (live) {
                off = p.offsets[p.u0 + ul];
                const int T = (int)(p.offsets[p.u0 + ul + 1] - off);
                Te = (p.first_frames > 0 && p.first_frames < T) ? p.first_frames : T;
            }
            // tile length = max Te over the 128 rows: every worker thread needs it (barrier counts);
            // named barrier over the 256 workers only (the MMA warp does not take part)
            if (tid == 0) s_Tt[it & 1] = 0;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (g == 0) atomicMax(&s_Tt[it & 1], Te);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int Tt = s_Tt[it & 1];

            float V[MG][8], Vx[MG];
            double base[MG];
#pragma unroll
            for (int k = 0; k < MG; k++) {
                Vx[k] = -INFINITY; base[k] = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) V[k][j] = -INFINITY;
            }
            // raw feature prefetch registers: up to 8 chunks per thread
            float4 xr[8];
            auto load_row = [&](int t) {
                const bool ok = live && t < Te;
                const float4 *row = reinterpret_cast<const float4 *>(p.X + (size_t)(off + t) * p.ldx);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int ch = cbeg + c;
                    xr[c] = (ok && ch < cend && 4 * ch < p.ldx) ? __ldg(row + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto convert_row = [&](uint32_t fr) {      // write this thread's chunks of frame fr into stage fr & 1
                unsigned char *aS = sA + (fr & 1) * a_stage + (size_t)(r >> 3) * (2 * nck * 128) + (r & 7) * 16;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int ch = cbeg + c;
                    if (ch < cend) {
                        const float xv[4] = {xr[c].x, xr[c].y, xr[c].z, xr[c].w};
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int d = 4 * ch + q;
                            const float2 gsd = sGs[d];
                            float xs = (xv[q] - gsd.x) * gsd.y;
                            xs = fminf(fmaxf(xs, -250.f), 250.f);
                            if (d >= p.D) xs = (d == p.D) ? 1.f : 0.f;      // constant slot / zero padding
                            const float x2 = xs * xs;
                            const __half2 h = __floats2half2_rn(xs, x2);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(xs - hf.x, x2 - hf.y);
                            hi[q] = *reinterpret_cast<const uint32_t *>(&h);
                            lo[q] = *reinterpret_cast<const uint32_t *>(&l);
                        }
                        *reinterpret_cast<uint4 *>(aS + (size_t)ch * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(aS + (size_t)(nck + ch) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async();
                mbar_arrive(barA_full + 8 * (fr & 1));
            };

            if (Tt > 0) {
                load_row(0);
                convert_row(f);
                load_row(1);
            }
            uint16_t *bpp = p.bp + ul;
            for (int t = 0; t < Tt; t++, f++) {
                if (t + 1 < Tt) {
                    convert_row(f + 1);
                    load_row(t + 2);
                }
                const uint32_t s = f & 1, ph = (f >> 1) & 1;
                mbar_wait(barAcc_full + 8 * s, ph);
                tc_fence_after();
                uint32_t ev[MG][8];
#pragma unroll
                for (int k = 0; k < MG; k++)
                    if (k < mcnt) tmem_ld8(tmem_lane + s * (uint32_t)ncols + (uint32_t)(mbeg + k) * 8, ev[k]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(barAcc_empty + 8 * s);
                if (p.dbgE && live && t < Te) {
#pragma unroll
                    for (int k = 0; k < MG; k++)
                        if (k < mcnt)
#pragma unroll
                            for (int j = 0; j < 8; j++) p.dbgE[(size_t)(off + t) * ncols + (mbeg + k) * 8 + j] = __uint_as_float(ev[k][j]);
                }
                if (live && t < Te) {
#pragma unroll
                    for (int k = 0; k < MG; k++) {
                        if (k < mcnt) {
                            const float4 *tr = sTr + (size_t)(mbeg + k) * 5;
                            const float4 t0 = tr[0], t1 = tr[1], t2 = tr[2], t3 = tr[3], t4 = tr[4];
                            const float stay[8] = {t0.x, t0.z, t1.x, t1.z, t2.x, t2.z, t3.x, t3.z};
                            const float adv[8] = {t0.y, t0.w, t1.y, t1.w, t2.y, t2.w, t3.y, t3.w};
                            float e[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) e[j] = __uint_as_float(ev[k][j]);
                            if (t == 0) {
                                V[k][0] = t4.z + e[0];                       // V[0,1] = ln A01 + E[0,1]
                            } else {
                                uint32_t sb = 0;                             // "stayed" bits, exit first
                                float nx = -INFINITY;
                                if (t >= 8) {                                // exit opens at t >= N (custom_hmm.py:481-485)
                                    const float ca = V[k][7] + t4.y, cs = Vx[k] + t4.x;
                                    sb = __funnelshift_l(__float_as_uint(ca - cs), sb, 1);
                                    nx = fmaxf(ca, cs);
                                } else {
                                    sb = 1;
                                }
#pragma unroll
                                for (int j = 7; j >= 1; j--) {
                                    const float ca = V[k][j - 1] + adv[j], cs = V[k][j] + stay[j];
                                    sb = __funnelshift_l(__float_as_uint(ca - cs), sb, 1);   // sign(ca - cs) = 1 -> stayed
                                    V[k][j] = fmaxf(ca, cs) + e[j];
                                }
                                {
                                    const float ca = (t == 1) ? t4.z : -INFINITY;            // entry only at t == 1
                                    const float cs = V[k][0] + stay[0];
                                    sb = __funnelshift_l(__float_as_uint(ca - cs), sb, 1);
                                    V[k][0] = fmaxf(ca, cs) + e[0];
                                }
                                Vx[k] = nx;
                                // after 9 shifts: bit 8 = exit, bit j = state j+1; advanced = !stayed
                                bpp[((size_t)(mbeg + k) * p.maxT + t) * p.Bpad] = (uint16_t)((~sb) & 0x1FFu);
                                if ((t & 3) == 0) {                          // renormalise, offset kept in float64
                                    float mx = Vx[k];
#pragma unroll
                                    for (int j = 0; j < 8; j++) mx = fmaxf(mx, V[k][j]);
                                    if (mx > -INFINITY && mx < INFINITY) {
#pragma unroll
                                        for (int j = 0; j < 8; j++) V[k][j] -= mx;
                                        Vx[k] -= mx;
                                        base[k] += (double)mx;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (live) {
#pragma unroll
                for (int k = 0; k < MG; k++)
                    if (k < mcnt) {
                        const double sc = (Te > 0 && Vx[k] > -INFINITY) ? (double)Vx[k] + base[k] : -INFINITY;
                        p.scores[(size_t)ul * M + mbeg + k] = sc;
                    }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, tcols);
}

// ------------------------------------------------------------------------------------------------
bool sapr_tc_eligible(const sapr_models *m) {
    return m->emission == SAPR_EMIT_DIAG && m->topology == SAPR_TOPO_ENTRY_EXIT && m->N == 8 && m->M <= 12 &&
           m->D + 1 <= 64;
}

size_t sapr_tc_image_bytes(const sapr_models *m, int *nck_out, int *ncols_out) {
    int nck = (m->D + 1 + 3) / 4;
    nck = (nck + 1) / 2 * 2;
    const int ncols = (m->M * 8 + 15) / 16 * 16;
    if (nck_out) *nck_out = nck;
    if (ncols_out) *ncols_out = ncols;
    size_t w = (size_t)2 * (ncols / 8) * nck * 64 * sizeof(__half);
    w = (w + 255) / 256 * 256;
    size_t g = ((size_t)4 * nck * sizeof(float2) + 255) / 256 * 256;
    size_t t = ((size_t)m->M * 5 * sizeof(float4) + 255) / 256 * 256;
    return w + g + t;
}

int sapr_tc_prepare(sapr_models *m) {
    sapr_ctx *ctx = m->ctx;
    int nck, ncols;
    sapr_tc_image_bytes(m, &nck, &ncols);
    const size_t w = ((size_t)2 * (ncols / 8) * nck * 64 * sizeof(__half) + 255) / 256 * 256;
    const size_t g = ((size_t)4 * nck * sizeof(float2) + 255) / 256 * 256;
    __half *wimg = (__half *)m->tc_image;
    float2 *gs = (float2 *)((char *)m->tc_image + w);
    float4 *trp = (float4 *)((char *)m->tc_image + w + g);
    {
        const size_t smem = (size_t)2 * (ncols / 8) * nck * 128 + (size_t)2 * 16 * 2 * nck * 128 + (size_t)m->M * 5 * 16 +
                            (size_t)4 * nck * 8 + 16 + 64 + 16;
        const int MG = (m->M + 1) / 2;
        int nb = 1;
        cudaError_t e;
        if (MG <= 2) { cudaFuncSetAttribute(k_viterbi_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                       e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_viterbi_tc<2>, TC_THREADS, smem); }
        else if (MG <= 4) { cudaFuncSetAttribute(k_viterbi_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_viterbi_tc<4>, TC_THREADS, smem); }
        else { cudaFuncSetAttribute(k_viterbi_tc<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
               e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_viterbi_tc<6>, TC_THREADS, smem); }
        if (e != cudaSuccess) { cudaGetLastError(); nb = 1; }
        int tcols = 32;
        while (tcols < 2 * ncols) tcols <<= 1;
        nb = std::max(1, std::min(nb, 512 / tcols));     // TMEM: 512 columns per SM
        m->tc_ctas_per_sm = nb;
    }
    k_prepare_tc<<<1, 256, sizeof(double) * 2 * 4 * nck, ctx->stream>>>(m->M, m->S, m->D, nck, ncols, m->mean, m->cov, m->la64,
                                                                         m->lb64, wimg, gs, trp);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

// argmax + back-trace launcher shared with the SIMT path (viterbi.cu)
int sapr_viterbi_finish_u16(sapr_ctx *ctx, const int64_t *offsets, int u0, int nu, int N, int nslots, int first_frames,
                            const uint16_t *bp, int64_t Bpad, int maxT, const double *scores, int32_t *best_word,
                            double *best_score, double *scores_out, int M, uint8_t *best_path, uint8_t *all_paths,
                            int64_t total_frames);

int sapr_viterbi_tc_launch(sapr_ctx *ctx, sapr_models *m, const float *X, int ldx, const int64_t *offsets, int B,
                           int64_t total_frames, int max_T, int first_frames, int32_t *best_word, double *best_score,
                           double *scores, uint8_t *best_path, uint8_t *all_paths, float *dbgE) {
    int nck, ncols;
    sapr_tc_image_bytes(m, &nck, &ncols);
    const size_t w = ((size_t)2 * (ncols / 8) * nck * 64 * sizeof(__half) + 255) / 256 * 256;
    const size_t g = ((size_t)4 * nck * sizeof(float2) + 255) / 256 * 256;
    const int M = m->M;
    const int Tm = (first_frames > 0 && first_frames < max_T) ? first_frames : max_T;
    int64_t per_utt = (int64_t)M * (Tm > 0 ? Tm : 1) * sizeof(uint16_t);
    int chunk = (int)std::min<int64_t>(B, std::max<int64_t>(TC_ROWS, ((int64_t)1024 << 20) / per_utt));
    chunk = (chunk + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
    const int64_t Bpad = chunk;
    int rc = sapr_ws_reserve(ctx, 0, (size_t)per_utt * Bpad);
    if (rc) return rc;
    if ((rc = sapr_ws_reserve(ctx, 1, (size_t)chunk * M * sizeof(double)))) return rc;
    uint16_t *bp = (uint16_t *)ctx->ws[0];
    double *sc_ws = (double *)ctx->ws[1];
    const size_t smem = (size_t)2 * (ncols / 8) * nck * 128 + (size_t)2 * 16 * 2 * nck * 128 + (size_t)M * 5 * 16 +
                        (size_t)4 * nck * 8 + 16 + 64 + 16;
    const int MG = (M + 1) / 2;
    auto launch = [&](auto kern, const TcParams &prm, int grid) -> int {
        SAPR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            ProfScope ps(ctx, 0);
            kern<<<grid, TC_THREADS, smem, ctx->stream>>>(prm);
        }
        SAPR_LAUNCH_CHECK(ctx);
        return SAPR_OK;
    };
    for (int u0 = 0; u0 < B; u0 += chunk) {
        const int nu = std::min(chunk, B - u0);
        TcParams prm;
        prm.X = X; prm.ldx = ldx; prm.offsets = offsets; prm.u0 = u0; prm.nu = nu; prm.M = M; prm.D = m->D; prm.nck = nck;
        prm.ncols = ncols; prm.first_frames = first_frames; prm.wimg = (const __half *)m->tc_image;
        prm.gs = (const float2 *)((const char *)m->tc_image + w); prm.trp = (const float4 *)((const char *)m->tc_image + w + g);
        prm.bp = bp; prm.Bpad = Bpad; prm.maxT = Tm; prm.scores = sc_ws; prm.dbgE = dbgE;
        const int ntiles = (nu + TC_ROWS - 1) / TC_ROWS;
        const int grid = std::min(ntiles, ctx->sm_count * m->tc_ctas_per_sm);
        if (MG <= 2) rc = launch(k_viterbi_tc<2>, prm, grid);
        else if (MG <= 4) rc = launch(k_viterbi_tc<4>, prm, grid);
        else rc = launch(k_viterbi_tc<6>, prm, grid);
        if (rc) return rc;
        if ((rc = sapr_viterbi_finish_u16(ctx, offsets, u0, nu, m->N, M, first_frames, bp, Bpad, Tm, sc_ws, best_word, best_score,
                                          scores, M, best_path, all_paths, total_frames)))
            return rc;
    }
    return SAPR_OK;
}
