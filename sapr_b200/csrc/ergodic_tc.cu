// ergodic_tc.cu -- forward (score) pass of a large fully connected HMM on the tensor cores (BASELINE cfg 4: N = 256
// all-emitting states, dense transition matrix, diagonal-Gaussian emission).
//
// Replaces hmmlearn's GaussianHMM.score for the model hmmlearn_hmm.py:27-43 builds, at the state counts where the
// transition step is a real dense contraction (forward_log; hmmlearn is not vendored, parity is
// against this repo's float64 restatement).  fp32 production mode; the float64 path of hmmlearn.cu is the verification mode.
//
// Scaled linear-domain forward recursion, 128 utterances per CTA at the same frame index:
//     v_t(j) = kappa_t b_t(j) * sum_i alpha^_{t-1}(i) A(i, j),   b_t(j) = exp(lf_t(j) - m_t),  m_t = max_j lf_t(j)
//     alpha^_t = v_t,  kappa_{t+1} = C / sum_j v_t(j)  (the normaliser lags one frame, so the products go straight from the
//     accumulator into the next A operand; a frame whose mass collapses is rewritten with its exact normaliser),
//     log P = sum_t (m_t - ln kappa_t) + ln sum_j alpha^_{T-1}(j)
// The contraction over i is one [128 x S] . [S x S] product per frame on tcgen05: the A operand (alpha^, fp16 hi/lo
// split, 22 bits) is written by the threads that own the rows straight into TMEM, the transition matrix (fp16, S*S*2
// bytes = 128 KB at S = 256) stays resident in shared memory, the fp32 accumulator [128 x S] lives in TMEM and
// tcgen05.ld hands thread (row r, quarter g) its S/4 states.  While the tensor pipe runs frame t's product, the worker
// threads load frame t's emissions and turn them into kappa_t b_t(j) in registers.
// Emissions are a second tensor-core contraction, [x', x'^2, 1] . W_e (k_erg_emission_tc, the operand construction of
// viterbi_tc.cu with S columns), written to HBM as fp32 in the layout the forward kernel reads coalesced:
//     lf[tile][t][j/4][row][4],  rmax[tile][t][quarter][row]     (row = utterance within the 128-utterance tile)
// 2 * S*4 bytes per utterance-frame (write + read) is the HBM traffic of the pair and its bound: TMEM cannot hold the
// emission and the transition accumulator of 256 states at once (2 x 256 columns + both A operands > 512 columns), so the two
// contractions are two kernels.
#include "tc_common.cuh"

#define ERG_MAX_S 256
#define ERG_WS_BYTES ((size_t)24 << 30)   /* emission staging per chunk of tiles */
#define ERG_W_SCALE 32768.0        /* weight image = 2^15 A: fp16 then spans transition probabilities from 1 down to 2^-39 (normal above 2^-29) */
#define ERG_ALPHA_SCALE 32768.0f   /* stored vector = 32768 alpha^ (sum over states): fp16 hi/lo parts stay normal */
#define ERG_RESCUE (1.0f / 4194304.0f)  /* redo a frame exactly when its lagged-normalised mass falls below this: the dominant entry keeps >= 15 bits */

// ------------------------------------------------------------------------------------------------
// transition image: W[n = j][k = i] = 256 A[i][j] as fp16, K-major no-swizzle core matrices (8 rows x 8 halves)
__global__ void k_erg_prepare(int S, const double *__restrict__ A, const double *__restrict__ pi, __half *__restrict__ wimg,
                              float *__restrict__ pif) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < S * S) {
        const int n = idx / S, k = idx % S;                       // n = destination state j, k = source state i
        const size_t o = ((size_t)(n / 8) * (S / 8) + k / 8) * 64 + (size_t)(n % 8) * 8 + (k % 8);
        wimg[o] = __double2half(ERG_W_SCALE * A[(size_t)k * S + n]);
    }
    if (idx < S) pif[idx] = (float)pi[idx];
}

// emission image (same construction as k_prepare_tc of viterbi_tc.cu, one column per state of the dense model):
//   eimg[2][S/8][nck][8][8] halves (hi plane, lo plane); a chunk is 4 feature dims = 8 K elements [x'_0..3, x'^2_0..3];
//   x' = x s + b standardises with the centre / spread of the state means; dim index D is the constant slot (x' = 1).
//   sb[2][4*nck] float (s, b).
__global__ void k_erg_prepare_emis(int S, int D, int nck, const double *__restrict__ mean, const double *__restrict__ var,
                                   __half *__restrict__ eimg, float *__restrict__ sb) {
    extern __shared__ double s_gs[];   // [2][4*nck]: centre g, scale s
    const int nd = 4 * nck;
    // one warp per feature dim: lanes stride over the states, fixed-order shuffle tree (a single thread walking all S states
    // per dim took 0.47 ms at S = 256 -- 4 % of a scoring call)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    auto wsum = [](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    for (int d = warp; d < nd; d += nwarp) {
        double g = 0.0, sc = 0.0;
        float sf = 0.f, bf = 0.f;
        if (d < D) {
            double sm = 0.0;
            for (int j = lane; j < S; j += 32) sm += mean[(size_t)j * D + d];
            g = wsum(sm) / S;
            double v = 0.0;
            for (int j = lane; j < S; j += 32) {
                const double df = mean[(size_t)j * D + d] - g;
                v += var[(size_t)j * D + d] + df * df;
            }
            v = wsum(v) / S;
            sc = (v > 0 && v < 1e300) ? 4.0 / sqrt(v) : 1.0;
            sf = (float)sc; bf = (float)(-g * sc);
            sc = (double)sf; g = -(double)bf / sc;   // the kernel standardises in fp32 with exactly (sf, bf)
        } else if (d == D) {
            sf = 0.f; bf = 1.f;
        }
        if (lane == 0) {
            s_gs[d] = g; s_gs[nd + d] = sc;
            if (blockIdx.x == 0) { sb[d] = sf; sb[nd + d] = bf; }       // every CTA recomputes the (cheap) constants, one writes them
        }
    }
    __syncthreads();
    const size_t plane = (size_t)(S / 8) * nck * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < S * nd; idx += gridDim.x * blockDim.x) {
        const int n = idx / nd, d = idx % nd;
        const double *mu = mean + (size_t)n * D, *vr = var + (size_t)n * D;
        double wx = 0.0, wx2 = 0.0;
        if (d < D) {
            const double g = s_gs[d], sc = s_gs[nd + d];
            const double hp = 0.5 / vr[d] / (sc * sc), mp = (mu[d] - g) * sc;
            wx = 2.0 * mp * hp; wx2 = -hp;
        } else if (d == D) {
            double ld = 0.0, c2 = 0.0;
            for (int q = 0; q < D; q++) {
                const double g = s_gs[q], sc = s_gs[nd + q];
                const double hp = 0.5 / vr[q] / (sc * sc), mp = (mu[q] - g) * sc;
                ld += log(vr[q]); c2 += mp * mp * hp;
            }
            const double cst = -0.5 * (D * SAPR_LOG2PI + ld) - c2;
            const __half ch = __double2half(cst);
            const double r1 = cst - (double)__half2float(ch);
            const __half cl = __double2half(r1);
            wx = cst; wx2 = r1 - (double)__half2float(cl);   // third-order piece of the constant rides on x'^2 = 1
        }
        const size_t o = ((size_t)(n / 8) * nck + d / 4) * 64 + (size_t)(n % 8) * 8 + (d % 4);
        const __half hx = __double2half(wx), hx2 = __double2half(wx2);
        eimg[o] = hx; eimg[o + 4] = hx2;
        eimg[plane + o] = __double2half(wx - (double)__half2float(hx));
        eimg[plane + o + 4] = __double2half(wx2 - (double)__half2float(hx2));
    }
}

// ------------------------------------------------------------------------------------------------
struct ErgParams {
    const float *X; int ldx; const int64_t *offsets; int B, S, D, nck, maxT, tile0, ntiles;   // tiles [tile0, tile0 + ntiles)
    const __half *wimg, *eimg; const float *pif, *sb;
    float *lf, *rmax;           // lf[ntiles][maxT][S/4][128][4], rmax[ntiles][maxT][4][128]
    double *logprob;
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

#define ERG_THREADS (TC_WORKERS + 32)
#define ERG_FB 16         /* frames per work item of the emission kernel */

__device__ __forceinline__ int erg_row_frames(const ErgParams &p, int u, int64_t &off) {
    off = 0;
    if (u >= p.B) return 0;
    off = p.offsets[u];
    return (int)(p.offsets[u + 1] - off);
}
__device__ __forceinline__ int erg_tile_frames(const ErgParams &p, int tile, int lane) {   // longest utterance of the tile (warp-wide)
    int Tt = 0;
    for (int r = lane; r < TC_ROWS; r += 32) { int64_t o; Tt = max(Tt, erg_row_frames(p, tile * TC_ROWS + r, o)); }
    for (int o = 16; o > 0; o >>= 1) Tt = max(Tt, __shfl_xor_sync(0xffffffffu, Tt, o));
    return min(Tt, p.maxT);
}

// ------------------------------------------------------------------------------------------------
// emission: work item = (tile, block of ERG_FB frames); per frame one [128 x 8 nck] . [8 nck x S] product in three fp16
// passes (lo Whi, hi Wlo, hi Whi), A operand written to TMEM by the threads that own the rows (2 stages), fp32 accumulator
// [128 x S] in TMEM (single: the epilogue of frame k precedes the product of frame k + 1, the conversion of k + 2 overlaps it).
__global__ void __launch_bounds__(ERG_THREADS, 1) k_erg_emission_tc(const ErgParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int S = p.S, nck = p.nck, nd = 4 * nck;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t plane_bytes = (uint32_t)(S / 8) * nck * 128u;
    unsigned char *sW = smem;
    float *sSb = reinterpret_cast<float *>(smem + 2 * plane_bytes);            // [2][nd]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sSb + 2 * nd);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + 4);
    const uint32_t barA_full = smem_u32(sBar), barAcc_full = barA_full + 16, barAcc_empty = barA_full + 24;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.eimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        for (uint32_t i = tid; i < 2 * plane_bytes / 16; i += ERG_THREADS) dst[i] = src[i];
        for (int i = tid; i < 2 * nd; i += ERG_THREADS) sSb[i] = p.sb[i];
    }
    if (tid == 0) {
        mbar_init(barA_full, TC_WORKER_WARPS);
        mbar_init(barA_full + 8, TC_WORKER_WARPS);
        mbar_init(barAcc_full, 1);
        mbar_init(barAcc_empty, TC_WORKER_WARPS);
        fence_barrier_init();
    }
    if (warp == TC_WORKER_WARPS) tmem_alloc(smem_u32(sTmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const uint32_t tmem_acc = tmem_base, tmem_a0 = tmem_base + (uint32_t)S, a_stage_cols = 8u * nck;
    const int nblk = (p.maxT + ERG_FB - 1) / ERG_FB, nitems = p.ntiles * nblk;
    uint32_t k = 0;       // frames processed by this CTA so far (selects A stage and barrier parities)

    if (warp == TC_WORKER_WARPS) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(S >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        const uint64_t dHi = make_desc(smem_u32(sW), 128, (uint32_t)nck * 128u);
        const uint64_t dLo = make_desc(smem_u32(sW) + plane_bytes, 128, (uint32_t)nck * 128u);
        const int nks = nck / 2;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tl = item / nblk, t0 = (item % nblk) * ERG_FB;
            const int t1 = min(t0 + ERG_FB, erg_tile_frames(p, p.tile0 + tl, lane));
            for (int t = t0; t < t1; t++, k++) {
                mbar_wait(barA_full + 8 * (k & 1u), (k >> 1) & 1u);
                if (k > 0) mbar_wait(barAcc_empty, (k - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t ahi = tmem_a0 + (k & 1u) * a_stage_cols, alo = ahi + 4u * nck;
                    for (int ks = 0; ks < nks; ks++) umma_f16_ts(tmem_acc, alo + 8 * ks, dHi + 16 * ks, idesc, ks > 0);
                    for (int ks = 0; ks < nks; ks++) umma_f16_ts(tmem_acc, ahi + 8 * ks, dLo + 16 * ks, idesc, 1);
                    for (int ks = 0; ks < nks; ks++) umma_f16_ts(tmem_acc, ahi + 8 * ks, dHi + 16 * ks, idesc, 1);
                    umma_commit(barAcc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== workers: thread (row r, quarter g) =====================
        const int q = warp & 3, g = warp >> 2, r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const int spt = S / TC_GROUPS, j0 = g * spt, nch = spt / 16;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tl = item / nblk, t0 = (item % nblk) * ERG_FB;
            const int t1 = min(t0 + ERG_FB, erg_tile_frames(p, p.tile0 + tl, lane));
            int64_t off;
            const int T = erg_row_frames(p, (p.tile0 + tl) * TC_ROWS + r, off);
            // conversion of frame t into A stage (kk & 1): chunks g, g + 4, g + 8 of this row
            auto convert = [&](int t, uint32_t kk) {
                const uint32_t ahi = tmem_a0 + (kk & 1u) * a_stage_cols + lane_sel, alo = ahi + 4u * nck;
                const float *xr = p.X + (size_t)(off + t) * p.ldx;
                const bool act = t < T;
                for (int c = g; c < nck; c += TC_GROUPS) {
                    const float4 sc = *reinterpret_cast<const float4 *>(sSb + 4 * c);
                    const float4 bc = *reinterpret_cast<const float4 *>(sSb + nd + 4 * c);
                    float x[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = (act && 4 * c + i < p.D) ? __ldg(xr + 4 * c + i) : 0.f;
                    const float2 a01 = make_float2(fmaf(x[0], sc.x, bc.x), fmaf(x[1], sc.y, bc.y));
                    const float2 a23 = make_float2(fmaf(x[2], sc.z, bc.z), fmaf(x[3], sc.w, bc.w));
                    const float2 q01 = mul2(a01, a01), q23 = mul2(a23, a23);
                    const uint32_t h0 = pack_h2(a01), h1 = pack_h2(a23), h2 = pack_h2(q01), h3 = pack_h2(q23);
                    tmem_st4(ahi + 4 * c, h0, h1, h2, h3);
                    tmem_st4(alo + 4 * c, pack_h2(sub2(a01, unpack_h2(h0))), pack_h2(sub2(a23, unpack_h2(h1))),
                             pack_h2(sub2(q01, unpack_h2(h2))), pack_h2(sub2(q23, unpack_h2(h3))));
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(barA_full + 8 * (kk & 1u));
            };
            if (t0 < t1) convert(t0, k);
            float *lft = p.lf + (size_t)tl * p.maxT * S * TC_ROWS;
            float *rmt = p.rmax + (size_t)tl * p.maxT * TC_GROUPS * TC_ROWS;
            for (int t = t0; t < t1; t++, k++) {
                if (t + 1 < t1) convert(t + 1, k + 1);
                mbar_wait(barAcc_full, k & 1u);
                tc_fence_after();
                float mx = -INFINITY;
                float4 *dst = reinterpret_cast<float4 *>(lft) + ((size_t)t * (S / 4) + j0 / 4) * TC_ROWS + r;
                for (int c = 0; c < nch; c++) {
                    uint32_t ev[16];
                    tmem_ld16(tmem_acc + lane_sel + (uint32_t)(j0 + 16 * c), ev);
                    tmem_ld_wait();
#pragma unroll
                    for (int i4 = 0; i4 < 4; i4++) {
                        const float4 v = make_float4(__uint_as_float(ev[4 * i4]), __uint_as_float(ev[4 * i4 + 1]),
                                                     __uint_as_float(ev[4 * i4 + 2]), __uint_as_float(ev[4 * i4 + 3]));
                        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                        dst[(size_t)(4 * c + i4) * TC_ROWS] = v;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(barAcc_empty);
                rmt[((size_t)t * TC_GROUPS + g) * TC_ROWS + r] = mx;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WORKER_WARPS) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// forward recursion over the emissions of k_erg_emission_tc
template <int NCH>   // 16-state chunks per thread = S / 64
__global__ void __launch_bounds__(ERG_THREADS, 1) k_erg_forward_tc(const ErgParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int S = 64 * NCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w_bytes = (uint32_t)S * S * 2u;
    unsigned char *sW = smem;
    float *sSum = reinterpret_cast<float *>(smem + w_bytes);                    // [2][128][4] partial row sums (frame parity) + [128][4] rare path
    float *sPi = sSum + 3 * TC_ROWS * TC_GROUPS;                               // [S]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sPi + S);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + 2);
    const uint32_t barA_full = smem_u32(sBar), barAcc_full = barA_full + 8;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        for (uint32_t i = tid; i < w_bytes / 16; i += ERG_THREADS) dst[i] = src[i];
        for (int i = tid; i < S; i += ERG_THREADS) sPi[i] = p.pif[i];
    }
    if (tid == 0) {
        mbar_init(barA_full, TC_WORKER_WARPS);
        mbar_init(barAcc_full, 1);
        fence_barrier_init();
    }
    if (warp == TC_WORKER_WARPS) tmem_alloc(smem_u32(sTmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const uint32_t tmem_acc = tmem_base, tmem_ahi = tmem_base + (uint32_t)S, tmem_alo = tmem_ahi + (uint32_t)S / 2u;
    uint32_t fa = 0;      // frames processed by this CTA so far: parity of A_full and of the partial-sum buffer
    uint32_t fc = 0;      // products committed so far: parity of acc_full (one per frame t >= 1)
    if (warp == TC_WORKER_WARPS) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(S >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        const uint64_t dW = make_desc(smem_u32(sW), 128, (uint32_t)(S / 8) * 128u);
        const int nks = S / 16;
        for (int tl = blockIdx.x; tl < p.ntiles; tl += gridDim.x) {
            const int Tt = erg_tile_frames(p, p.tile0 + tl, lane);
            for (int t = 0; t < Tt; t++, fa++) {
                mbar_wait(barA_full, fa & 1u);                    // alpha^_t stored by all worker warps
                if (t + 1 >= Tt) continue;                        // the last frame's vector feeds no product
                tc_fence_after();
                if (elect_one()) {
                    uint64_t d = dW;
                    uint32_t a = tmem_alo;
                    for (int ks = 0; ks < nks; ks++, a += 8, d += 16) umma_f16_ts(tmem_acc, a, d, idesc, ks > 0);   // lo part first
                    d = dW; a = tmem_ahi;
                    for (int ks = 0; ks < nks; ks++, a += 8, d += 16) umma_f16_ts(tmem_acc, a, d, idesc, 1);
                    umma_commit(barAcc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== workers: thread (row r, quarter g) owns S/4 states =====================
        const int q = warp & 3, g = warp >> 2, r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        constexpr int spt = S / TC_GROUPS;                        // states per thread
        const int j0 = g * spt;
        for (int tl = blockIdx.x; tl < p.ntiles; tl += gridDim.x) {
            const int Tt = erg_tile_frames(p, p.tile0 + tl, lane);
            const int u = (p.tile0 + tl) * TC_ROWS + r;
            int64_t off;
            const int Tfull = erg_row_frames(p, u, off);
            const int T = min(Tfull, p.maxT);
            const float4 *lft = reinterpret_cast<const float4 *>(p.lf + (size_t)tl * p.maxT * S * TC_ROWS) + (size_t)(j0 / 4) * TC_ROWS + r;
            const float *rmt = p.rmax + (size_t)tl * p.maxT * TC_GROUPS * TC_ROWS + r;
            // a_t = Z_t * (stored vector): logz = ln Z_t;  kappa = scale applied to this frame's products
            double logz = 0.0;
            float kappa = ERG_ALPHA_SCALE, Vlast = 0.f;
            bool alive = true;
            for (int t = 0; t < Tt; t++, fa++) {
                const bool act = t < T;
                // ---- in the shadow of this frame's transition product: emissions -> bk_j = exp(lf_j - m) * kappa ----
                float bk[spt];
                float m = 0.f;
                if (act) {
                    const float *rm = rmt + (size_t)t * TC_GROUPS * TC_ROWS;
                    m = fmaxf(fmaxf(__ldg(rm), __ldg(rm + TC_ROWS)), fmaxf(__ldg(rm + 2 * TC_ROWS), __ldg(rm + 3 * TC_ROWS)));
                    const float4 *lfr = lft + (size_t)t * (S / 4) * TC_ROWS;
#pragma unroll
                    for (int i4 = 0; i4 < spt / 4; i4++) {
                        const float4 l = __ldg(lfr + (size_t)i4 * TC_ROWS);
                        bk[4 * i4] = l.x; bk[4 * i4 + 1] = l.y; bk[4 * i4 + 2] = l.z; bk[4 * i4 + 3] = l.w;
                    }
                    const float ml2 = m * 1.4426950408889634f;
#pragma unroll
                    for (int i = 0; i < spt; i++) {
                        float b;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(fmaf(bk[i], 1.4426950408889634f, -ml2)));
                        bk[i] = b * kappa;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < spt; i++) bk[i] = 0.f;
                }
                if (t >= 1) {
                    mbar_wait(barAcc_full, fc & 1u);
                    fc++;
                    tc_fence_after();
                }
                // ---- v_j = s_j * bk_j straight into the A operand (fp16 hi/lo), partial row sum ----
                auto emit = [&]() -> float {
                    float lsum = 0.f;
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        uint32_t ev[16];
                        if (t >= 1) {
                            tmem_ld16(tmem_acc + lane_sel + (uint32_t)(j0 + 16 * c), ev);
                            tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i++) ev[i] = __float_as_uint(sPi[j0 + 16 * c + i]);
                        }
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float2 v2 = make_float2(__uint_as_float(ev[2 * i]) * bk[16 * c + 2 * i],
                                                    __uint_as_float(ev[2 * i + 1]) * bk[16 * c + 2 * i + 1]);
                            lsum += v2.x + v2.y;
                            hi[i] = pack_h2(v2);
                            lo[i] = pack_h2(sub2(v2, unpack_h2(hi[i])));
                        }
                        tmem_st8(tmem_ahi + lane_sel + (uint32_t)(j0 + 16 * c) / 2u, hi);
                        tmem_st8(tmem_alo + lane_sel + (uint32_t)(j0 + 16 * c) / 2u, lo);
                    }
                    return lsum;
                };
                float *srow = sSum + ((fa & 1u) * TC_ROWS + r) * TC_GROUPS;
                srow[g] = emit();
                asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");     // the four warps that share these 32 rows
                const float4 ps = *reinterpret_cast<const float4 *>(srow);
                float V = (ps.x + ps.y) + (ps.z + ps.w);
                float lscale = (t >= 1 ? (float)ERG_W_SCALE : 1.f) * kappa;    // a_t = Z_{t-1} e^{m} / lscale * (stored vector)
                // The products carry the normaliser of the PREVIOUS frame and the global row maximum.  When a frame's mass
                // collapses below what the fp16 pair resolves (or underflows: the reachable states sit > 87 nats under the best
                // one), the quadrant redoes the frame in the log domain with its exact maximum and normaliser; the accumulator
                // is still intact.  Same rows and same V in the four warps of the quadrant, hence the same decision.
                const bool low = act && alive && V < ERG_ALPHA_SCALE * ERG_RESCUE;
                if (__any_sync(0xffffffffu, low)) {
                    const float4 *lfr = lft + (size_t)t * (S / 4) * TC_ROWS;
                    float wmax = -INFINITY;
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        uint32_t ev[16];
                        if (t >= 1) {
                            tmem_ld16(tmem_acc + lane_sel + (uint32_t)(j0 + 16 * c), ev);
                            tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i++) ev[i] = __float_as_uint(sPi[j0 + 16 * c + i]);
                        }
#pragma unroll
                        for (int i4 = 0; i4 < 4; i4++) {
                            const float4 l = act ? __ldg(lfr + (size_t)(4 * c + i4) * TC_ROWS) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const float w = (act && alive) ? lv[i] + logf(fmaxf(__uint_as_float(ev[4 * i4 + i]), 0.f)) : -INFINITY;
                                bk[16 * c + 4 * i4 + i] = w;
                                wmax = fmaxf(wmax, w);
                            }
                        }
                    }
                    float *xrow = sSum + (2 * TC_ROWS + r) * TC_GROUPS;      // exchange area of the rare path
                    xrow[g] = wmax;
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
                    const float4 pm = *reinterpret_cast<const float4 *>(xrow);
                    const float m2 = fmaxf(fmaxf(pm.x, pm.y), fmaxf(pm.z, pm.w));
                    float lsum = 0.f;
#pragma unroll
                    for (int i = 0; i < spt; i++) {
                        const float v = (m2 > -INFINITY) ? expf(bk[i] - m2) : 0.f;
                        bk[i] = v;
                        lsum += v;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");   // everyone has read the maxima
                    xrow[g] = lsum;
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
                    const float4 p2 = *reinterpret_cast<const float4 *>(xrow);
                    const float V2 = (p2.x + p2.y) + (p2.z + p2.w);
                    const float rho = V2 > 0.f ? ERG_ALPHA_SCALE / V2 : 0.f;
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float2 v2 = make_float2(bk[16 * c + 2 * i] * rho, bk[16 * c + 2 * i + 1] * rho);
                            hi[i] = pack_h2(v2);
                            lo[i] = pack_h2(sub2(v2, unpack_h2(hi[i])));
                        }
                        tmem_st8(tmem_ahi + lane_sel + (uint32_t)(j0 + 16 * c) / 2u, hi);
                        tmem_st8(tmem_alo + lane_sel + (uint32_t)(j0 + 16 * c) / 2u, lo);
                    }
                    if (act && alive) {
                        if (V2 > 0.f) {
                            V = ERG_ALPHA_SCALE;
                            m = m2;
                            lscale = (t >= 1 ? (float)ERG_W_SCALE : 1.f) * rho;
                        } else {
                            alive = false;                  // no state is reachable: probability zero
                        }
                    }
                }
                if (act && alive) {
                    logz += (double)m - (double)logf(lscale);
                    Vlast = V;
                    kappa = ERG_ALPHA_SCALE / ((float)ERG_W_SCALE * V);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(barA_full);
            }
            // log P = ln Z_{T-1} + ln sum_j (stored vector); an utterance longer than the promised max_T has no valid score
            if (g == 0 && u < p.B)
                p.logprob[u] = (Tfull > p.maxT) ? __longlong_as_double(0x7ff8000000000000LL) : (T <= 0 ? 0.0 : alive ? logz + (double)logf(Vlast) : (double)-INFINITY);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WORKER_WARPS) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// fp32 tensor-core score of B utterances (each at most max_T frames) against dense model mi, S in {64, 128, 192, 256}
extern "C" int sapr_ergodic_score(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, const int64_t *offsets,
                                  int B, int max_T, double *logprob) {
    if (!ctx || !m || !X || !offsets || !logprob) return SAPR_E_INVALID;
    if (!m->valid) SAPR_FAIL(ctx, SAPR_E_INVALID, "ergodic_score: model parameters not set");
    if (mi < 0 || mi >= m->M) SAPR_FAIL(ctx, SAPR_E_INVALID, "ergodic_score: model index out of range");
    if (m->topology != SAPR_TOPO_DENSE || m->emission != SAPR_EMIT_DIAG)
        SAPR_FAIL(ctx, SAPR_E_INVALID, "ergodic_score: needs DENSE topology + DIAG emission");
    const int S = m->S, D = m->D;
    const int nck = 2 * ((D + 1 + 7) / 8);
    if (S % 64 || S < 64 || S > ERG_MAX_S || nck > 10)
        SAPR_FAIL(ctx, SAPR_E_RANGE, "ergodic_score: the tensor-core path takes 64, 128, 192 or 256 states and D <= 39");
    if (B <= 0) return SAPR_OK;
    if (max_T <= 0) SAPR_FAIL(ctx, SAPR_E_INVALID, "ergodic_score: max_T must be positive");
    // emissions are staged through HBM per chunk of tiles: at most ERG_WS_BYTES of lf at a time
    const int ntiles = (B + TC_ROWS - 1) / TC_ROWS;
    const size_t tile_lf = (size_t)max_T * S * TC_ROWS * 4, tile_rm = (size_t)max_T * TC_GROUPS * TC_ROWS * 4;
    int chunk = (int)std::max<size_t>(1, ERG_WS_BYTES / tile_lf);
    if (chunk >= ctx->sm_count) chunk = chunk / ctx->sm_count * ctx->sm_count;
    chunk = std::min(chunk, ntiles);
    const size_t w_b = (size_t)S * S * 2, e_b = (size_t)2 * (S / 8) * nck * 128;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t tot = al(w_b) + al(e_b) + al((size_t)S * 4) + al((size_t)8 * nck * 4) + al(tile_lf * chunk) + al(tile_rm * chunk);
    int rc = sapr_ws_reserve(ctx, 6, tot);
    if (rc) return rc;
    char *ws = (char *)ctx->ws[6];
    __half *wimg = (__half *)ws; ws += al(w_b);
    __half *eimg = (__half *)ws; ws += al(e_b);
    float *pif = (float *)ws; ws += al((size_t)S * 4);
    float *sb = (float *)ws; ws += al((size_t)8 * nck * 4);
    float *lf = (float *)ws; ws += al(tile_lf * chunk);
    float *rmax = (float *)ws;
    k_erg_prepare<<<(S * S + 255) / 256, 256, 0, ctx->stream>>>(S, m->A + (size_t)mi * S * S, m->pi + (size_t)mi * S, wimg, pif);
    SAPR_LAUNCH_CHECK(ctx);
    k_erg_prepare_emis<<<std::max(1, std::min(64, (S * 4 * nck + 255) / 256)), 256, sizeof(double) * 8 * nck, ctx->stream>>>(S, D, nck, m->mean + (size_t)mi * S * D,
                                                                          m->cov + (size_t)mi * S * D, eimg, sb);
    SAPR_LAUNCH_CHECK(ctx);
    ErgParams prm;
    prm.X = X; prm.ldx = ldx; prm.offsets = offsets; prm.B = B; prm.S = S; prm.D = D; prm.nck = nck; prm.maxT = max_T;
    prm.wimg = wimg; prm.eimg = eimg; prm.pif = pif; prm.sb = sb; prm.lf = lf; prm.rmax = rmax; prm.logprob = logprob;
    const size_t smem_e = e_b + (size_t)8 * nck * 4 + 64;
    const size_t smem_f = w_b + (size_t)3 * TC_ROWS * TC_GROUPS * 4 + (size_t)S * 4 + 64;
    SAPR_CUDA(ctx, cudaFuncSetAttribute(k_erg_emission_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
    void (*fwd)(const ErgParams) = S == 64 ? k_erg_forward_tc<1> : S == 128 ? k_erg_forward_tc<2> : S == 192 ? k_erg_forward_tc<3> : k_erg_forward_tc<4>;
    SAPR_CUDA(ctx, cudaFuncSetAttribute(fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    const int nblk = (max_T + ERG_FB - 1) / ERG_FB;
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        prm.tile0 = t0; prm.ntiles = std::min(chunk, ntiles - t0);
        {
            ProfScope ps(ctx, 4);
            k_erg_emission_tc<<<std::min(prm.ntiles * nblk, ctx->sm_count), ERG_THREADS, smem_e, ctx->stream>>>(prm);
        }
        SAPR_LAUNCH_CHECK(ctx);
        {
            ProfScope ps(ctx, 5);
            fwd<<<std::min(prm.ntiles, ctx->sm_count), ERG_THREADS, smem_f, ctx->stream>>>(prm);
        }
        SAPR_LAUNCH_CHECK(ctx);
    }
    return SAPR_OK;
}
