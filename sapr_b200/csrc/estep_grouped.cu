// estep_grouped.cu -- Baum-Welch E-step with the sufficient statistics fused into the backward sweep (fp32 production).
//
// Replaces custom_hmm.py:417-439 (emission, forward, backward, gamma, xi, accumulators) AND the sums update_B consumes
// (:372-386) in ONE kernel for the training layout the reference itself uses: utterances grouped by word
// (train.py:94 `load_mfccs_by_word`), here additionally equal-length and contiguous (BASELINE cfg 3).  Against
// k_estep_tc + k_stats_diag8 (the general path, kept for ragged / unsorted batches) it
//   * never writes gamma, the emissions or a second copy of anything O(frames x dims) to HBM: the only scratch is
//     alpha-hat (32 B per utterance-frame), written in the forward sweep and read back in the backward sweep;
//   * re-streams the features in reverse frame order for the backward sweep, recomputes the emissions on the (idle)
//     tensor pipe, forms gamma_t in registers and accumulates  Gamma_t^T . [x', x'^2, 1]  of the whole tile with
//     tcgen05.mma -- the fp16 hi/lo feature image a frame is converted into is ONE shared-memory block that is read
//     twice: as the K-major A operand of the emission product (rows = utterances) and as the MN-major A operand of the
//     statistics product (rows = features, K = utterances), the posteriors being the MN-major B operand;
//   * is fed by TMA tensor-map tile loads (one cp.async.bulk.tensor.3d per frame of a tile, issued by one thread).
// HBM traffic: two feature reads + the alpha-hat round trip (L2 permitting) against the one feature read the
// roofline counts, instead of 3.2x.
//
// Tile = 128 consecutive utterances of ONE model (a model's last tile is partly filled; its surplus rows carry zero
// posteriors).  A CTA walks a tile's frames 0 .. T-1 (forward) and then T-1 .. 0 (backward) through one pipeline:
//   TMA thread   raw fp32 frame -> shared-memory ring
//   8 converter warps (row quadrant x chunk half): standardise, square, fp16 hi/lo split -> A image stage
//   MMA warp     emission: acc[128 x 16] = A_lo.[W_hi|W_lo] + A_hi.[W_hi|W_lo]  (K-major A from shared memory);
//                backward sweep only: stats[128 features x 16] += A_hi^T.[g_hi|g_lo] + A_lo^T.[g_hi|g_lo] (MN-major)
//   4 recursion warps (thread = utterance): log-space forward / backward recursion as in k_estep_tc (same arithmetic,
//                same reference quirks D4-D6), posteriors -> fp16 hi/lo B operand; every 8 frames they drain the
//                statistics accumulator from TMEM into fp32 registers (the tensor core truncates its fp32 accumulator,
//                so sums are kept short), per tile they store the 128 x 8 partial sums.
// 14 warps = 448 threads: 144 registers per thread, so the recursions keep their constants in registers.
#include "tc_common.cuh"

#define EG_REC_WARPS 4
#define EG_CONV_WARPS 8
#define EG_MMA_WARP (EG_REC_WARPS + EG_CONV_WARPS)
#define EG_TMA_WARP (EG_MMA_WARP + 1)
#define EG_THREADS (32 * (EG_TMA_WARP + 1))
#ifndef EG_T_STAGES
#define EG_T_STAGES 4         /* A-operand stages of the emission product in TMEM */
#endif
#define EG_I_STAGES 3         /* shared-memory image stages (statistics product, backward frames) */
#ifndef EG_ACC
#define EG_ACC 6              /* emission accumulator stages (16 TMEM columns each): deep, so the conversion -> product -> recursion hand-offs overlap */
#endif
#ifndef EG_GAM
#define EG_GAM 4              /* posterior operand stages */
#endif
#define EG_NBAR (2 * EG_MAX_RAW + 2 * EG_T_STAGES + EG_I_STAGES + 2 * EG_ACC + 2 * EG_GAM + 4)
#define EG_MAX_RAW 8
#define EG_GROUP 8            /* frames per statistics accumulation group */
#define EG_PF 3               /* backward sweep: L1 prefetch distance of the alpha-hat scratch, in frames */

struct EgParams {
    int T, B, M, D, nck, ntiles, nraw;
    const int32_t *tile_model, *tile_u0, *tile_rows;
    const __half *wimg; uint32_t w_plane_halves;      // model-set image (k_prepare_tc): [2][ncols/8][nck][64] halves
    const float *sb; const float4 *trp;
    float *scratch;                                   // alpha-hat: [grid][T][8][128]
    float *ustats;                                    // [B][24]: G | Xi | occ per utterance
    double *loglik;                                   // [B]
    float *tpart;                                     // [ntiles][128][8]: per-tile sums, row = feature row of the A image
    uint32_t rw;                                      // bytes per row of the raw ring
    int pfd;                                          // L2 prefetch distance in pipeline frames
    int flags;                                        // tuning experiments (SAPR_EG_EXP): 1 = statistics right behind the emission, 2 = recursions skip their arithmetic, 4 = converters skip theirs
    long long *trace;
};

struct EgSmem { uint32_t a, raw, w, gam, tr, sb, bar, total; };
__host__ __device__ inline EgSmem eg_smem_layout(int M, int nck, int nraw, uint32_t rw) {
    EgSmem L;
    const uint32_t stage = 2u * 16u * nck * 128u;     // hi plane + lo plane, [16 row groups][nck chunks][8 rows][8 halves]
    L.a = 0;
    L.raw = EG_I_STAGES * stage;
    L.w = L.raw + (uint32_t)nraw * TC_ROWS * rw;
    L.gam = L.w + 2u * (2u * nck * 128u);             // two W buffers (alternate tiles)
    L.tr = L.gam + (uint32_t)EG_GAM * 4096u;          // posterior operands [16 row groups][2][8 rows][8 halves]
    L.sb = L.tr + (uint32_t)M * TC_TRQ * 16;
    L.bar = (L.sb + (uint32_t)8 * nck * 4 + 15u) & ~15u;
    L.total = L.bar + EG_NBAR * 8 + 16;
    return L;
}

__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

// log(exp(x) + exp(y)), branch-free fp32 form (the same as k_estep_tc)
__device__ __forceinline__ float eg_lae(float x, float y) {
    const float m = fmaxf(x, y);
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * fabsf(x - y)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    r = fmaf(r, 0.6931471805599453f, m);
    return (m == -INFINITY) ? m : r;
}
__device__ __forceinline__ float eg_exp(float x) {
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(1.4426950408889634f * x));
    return t;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// ring cursor: stage index + phase bit, advanced without divisions
struct Ring {
    uint32_t s, ph;
    __device__ __forceinline__ void next(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

template <int NCH, bool EXP = false>   // NCH = chunks per converter half (nck / 2), 0 = run-time; EXP: the what-if flags are honoured (tuning builds)
__global__ void __launch_bounds__(EG_THREADS, 1) k_estep_grouped(const EgParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nck = p.nck, M = p.M, T = p.T, nraw = p.nraw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rw = p.rw, frame_bytes = TC_ROWS * rw;
    const EgSmem L = eg_smem_layout(M, nck, nraw, rw);
    const uint32_t rg_stride = (uint32_t)nck * 128u;            // bytes between 8-row groups of the A image
    const uint32_t a_plane = 16u * rg_stride, a_stage = 2u * a_plane;
    const uint32_t w_buf = 2u * nck * 128u;                     // one model's [W_hi | W_lo]
    const uint32_t sA = smem_u32(smem + L.a), sRaw = smem_u32(smem + L.raw), sW = smem_u32(smem + L.w), sG = smem_u32(smem + L.gam);
    const float4 *sTr = reinterpret_cast<const float4 *>(smem + L.tr);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem + L.bar);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + EG_NBAR);
    const uint32_t bRawFull = smem_u32(sBar), bRawEmpty = bRawFull + 8 * EG_MAX_RAW;
    const uint32_t bAFull = bRawEmpty + 8 * EG_MAX_RAW, bAFree = bAFull + 8 * EG_T_STAGES;
    const uint32_t bImgFree = bAFree + 8 * EG_T_STAGES;          // shared-memory image stages (backward frames only)
    const uint32_t bAccFull = bImgFree + 8 * EG_I_STAGES, bAccEmpty = bAccFull + 8 * EG_ACC;
    const uint32_t bGamFull = bAccEmpty + 8 * EG_ACC, bGamFree = bGamFull + 8 * EG_GAM;
    const uint32_t bStFull = bGamFree + 8 * EG_GAM, bStEmpty = bStFull + 16;

    {
        float4 *dtr = reinterpret_cast<float4 *>(smem + L.tr);
        for (int i = tid; i < M * TC_TRQ; i += EG_THREADS) dtr[i] = p.trp[i];
        float *dsb = reinterpret_cast<float *>(smem + L.sb);
        for (int i = tid; i < 8 * nck; i += EG_THREADS) dsb[i] = p.sb[i];
        // the statistics product reads 16 feature groups per row group; groups >= nck run into the following rows /
        // buffers: keep everything it can touch finite from the start (results of those rows are never read)
        uint4 *za = reinterpret_cast<uint4 *>(smem + L.a);
        for (uint32_t i = tid; i < (L.w - L.a) / 16; i += EG_THREADS) za[i] = make_uint4(0, 0, 0, 0);
        uint4 *zg = reinterpret_cast<uint4 *>(smem + L.gam);
        for (uint32_t i = tid; i < EG_GAM * 4096 / 16; i += EG_THREADS) zg[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int s = 0; s < nraw; s++) { mbar_init(bRawFull + 8 * s, 1); mbar_init(bRawEmpty + 8 * s, EG_CONV_WARPS); }
        for (int s = 0; s < EG_T_STAGES; s++) { mbar_init(bAFull + 8 * s, EG_CONV_WARPS); mbar_init(bAFree + 8 * s, 1); }
        for (int s = 0; s < EG_I_STAGES; s++) mbar_init(bImgFree + 8 * s, 1);
        for (int s = 0; s < EG_ACC; s++) { mbar_init(bAccFull + 8 * s, 1); mbar_init(bAccEmpty + 8 * s, EG_REC_WARPS); }
        for (int s = 0; s < EG_GAM; s++) { mbar_init(bGamFull + 8 * s, EG_REC_WARPS); mbar_init(bGamFree + 8 * s, 1); }
        for (int s = 0; s < 2; s++) { mbar_init(bStFull + 8 * s, 1); mbar_init(bStEmpty + 8 * s, EG_REC_WARPS); }
        fence_barrier_init();
    }
    const uint32_t a_cols = 8u * nck;                             // TMEM columns of one A-operand stage: per K step [hi 8 | lo 8]
    uint32_t tcols = 128;
    while (tcols < 16u * EG_ACC + 32u + EG_T_STAGES * a_cols) tcols <<= 1;
    if (warp == EG_MMA_WARP) tmem_alloc(smem_u32(sTmem), tcols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const uint32_t tmem_acc = tmem_base, tmem_st = tmem_base + 16u * EG_ACC;   // EG_ACC x 16 emission columns, 2 x 16 statistics columns
    const uint32_t tmem_a = tmem_st + 32u;                                      // EG_T_STAGES A-operand stages of the emission product
    const int npf = 2 * T;                                            // pipeline frames per tile: forward, then backward

    if (warp == EG_TMA_WARP) {
        // ===================== TMA producer: one thread =====================
        if (lane == 0) {
            tma_prefetch_desc(&tmap);
            Ring rr = {0, 0};
            // L2 prefetch cursor: runs p.pfd pipeline frames ahead of the loads (over tile boundaries), so the ring's loads
            // find their boxes in L2 -- the ring alone (nraw x 22 KB per SM) does not cover the HBM latency at this rate
            int ptile = blockIdx.x, pi = 0;
            auto prefetch_next = [&]() {
                if (ptile >= p.ntiles) return;
                const int t = pi < T ? pi : npf - 1 - pi;
                tma_prefetch_l2_3d(&tmap, 0, t, p.tile_u0[ptile]);
                if (++pi == npf) { pi = 0; ptile += gridDim.x; }
            };
            for (int k = 0; k < p.pfd; k++) prefetch_next();
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int u0 = p.tile_u0[tile];
                for (int i = 0; i < npf; i++) {
                    const int t = i < T ? i : npf - 1 - i;
                    if (p.pfd > 0) prefetch_next();
                    mbar_wait(bRawEmpty + 8 * rr.s, rr.ph ^ 1u);
                    if (EXP && (p.flags & 16)) { mbar_arrive(bRawFull + 8 * rr.s); rr.next((uint32_t)nraw); continue; }   // tuning experiment: no loads
                    mbar_arrive_tx(bRawFull + 8 * rr.s, frame_bytes);
                    tma_load_3d(sRaw + rr.s * frame_bytes, &tmap, 0, t, u0, bRawFull + 8 * rr.s);
                    rr.next((uint32_t)nraw);
                }
            }
        }
    } else if (warp == EG_MMA_WARP) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_k = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);   // M = 128, N = 16, K-major
        const uint32_t idesc_mn = idesc_k | (1u << 15) | (1u << 16);                                            // both operands MN-major
        const int nks = nck / 2;
        Ring ar = {0, 0}, ir = {0, 0}, cr = {0, 0}, gr = {0, 0}, sr = {0, 0};   // TMEM A stages, image stages, accumulator stages, posterior operands, statistics buffers
        int kt = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, kt++) {
            const int m = p.tile_model[tile];
            const uint32_t wb = sW + (uint32_t)(kt & 1) * w_buf;
            {   // this tile's [W_hi | W_lo] into the buffer the previous tile does not use
                const uint4 *g_hi = reinterpret_cast<const uint4 *>(p.wimg + (size_t)m * nck * 64);
                const uint4 *g_lo = reinterpret_cast<const uint4 *>(p.wimg + p.w_plane_halves + (size_t)m * nck * 64);
                uint4 *d = reinterpret_cast<uint4 *>(smem + L.w + (uint32_t)(kt & 1) * w_buf);
                const int n16 = nck * 8;                               // uint4 per plane
                for (int i = lane; i < n16; i += 32) { d[i] = g_hi[i]; d[n16 + i] = g_lo[i]; }
                fence_proxy_async();
                __syncwarp();
            }
            // Two product streams, issued greedily by this one warp: the emission of pipeline frame ei (needs its A image and
            // a free accumulator stage) runs ahead; the statistics of backward frame si (needs the posteriors the recursion
            // warps write, releases the A image) follow as soon as they can.  A fixed order would chain the two through the
            // converters (statistics -> free A stage -> conversion -> emission).
            int sf = 0;                                               // statistics frames issued in this tile
            int ei = 0, si = T + 1;                                   // next emission frame, next statistics frame (pipeline indices)
            uint32_t a_hist[4] = {0, 0, 0, 0};                        // A stage of pipeline frame k at [k & 3]
            while (si < npf) {
                bool did = false;
                if (ei < npf && mbar_test(bAFull + 8 * ar.s, ar.ph) && mbar_test(bAccEmpty + 8 * cr.s, cr.ph ^ 1u)) {
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem_acc + cr.s * 16u;
                        const uint32_t a_hi = tmem_a + ar.s * a_cols, a_lo = a_hi + 8u;      // K step ks: hi at +16 ks, lo at +16 ks + 8
                        if (NCH > 0) {
#pragma unroll
                            for (int ks = 0; ks < NCH; ks++) umma_f16_ts(d, a_lo + 16u * ks, make_desc(wb + 256u * ks, 128, rg_stride), idesc_k, ks > 0);
#pragma unroll
                            for (int ks = 0; ks < NCH; ks++) umma_f16_ts(d, a_hi + 16u * ks, make_desc(wb + 256u * ks, 128, rg_stride), idesc_k, 1u);
                        } else {
                            for (int ks = 0; ks < nks; ks++) umma_f16_ts(d, a_lo + 16u * ks, make_desc(wb + 256u * ks, 128, rg_stride), idesc_k, ks > 0);
                            for (int ks = 0; ks < nks; ks++) umma_f16_ts(d, a_hi + 16u * ks, make_desc(wb + 256u * ks, 128, rg_stride), idesc_k, 1u);
                        }
                        umma_commit(bAccFull + 8 * cr.s);
                        umma_commit(bAFree + 8 * ar.s);               // the TMEM operand is free once the emission product has read it
                    }
                    __syncwarp();
                    if (ei > T) { a_hist[ei & 3] = ir.s; ir.next(EG_I_STAGES); }   // backward frames with statistics also have a shared-memory image
                    ar.next(EG_T_STAGES);
                    cr.next(EG_ACC);
                    ei++;
                    did = true;
                }
                if (si < ei && mbar_test(bGamFull + 8 * gr.s, gr.ph) &&
                    (sf % EG_GROUP != 0 || mbar_test(bStEmpty + 8 * sr.s, sr.ph ^ 1u))) {
                    // Gamma^T . X of one frame: 8 K steps of 16 utterance rows, hi plane then lo plane
                    tc_fence_after();
                    const uint32_t astage = a_hist[si & 3];
                    const bool last = sf % EG_GROUP == EG_GROUP - 1 || sf == T - 2;
                    if (elect_one()) {
                        const uint32_t d = tmem_st + sr.s * 16u;
                        const uint32_t ab = sA + astage * a_stage, gb = sG + gr.s * 4096u;
                        if (!EXP || !(p.flags & 8)) {
#pragma unroll
                        for (int ks = 0; ks < 8; ks++)
                            umma_f16_ss(d, make_desc(ab + 2u * ks * rg_stride, rg_stride, 128), make_desc(gb + 2u * ks * 256u, 256, 128), idesc_mn,
                                        (sf % EG_GROUP != 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                        for (int ks = 0; ks < 8; ks++)
                            umma_f16_ss(d, make_desc(ab + a_plane + 2u * ks * rg_stride, rg_stride, 128), make_desc(gb + 2u * ks * 256u, 256, 128), idesc_mn, 1u);
                        }
                        umma_commit(bImgFree + 8 * astage);
                        umma_commit(bGamFree + 8 * gr.s);
                        if (last) umma_commit(bStFull + 8 * sr.s);
                    }
                    __syncwarp();
                    if (last) sr.next(2);
                    gr.next(EG_GAM);
                    sf++; si++;
                    did = true;
                }
                if (!did) {      // nothing ready: park on the barrier the next product of the leading stream waits for
                    if (ei < npf) mbar_try(bAFull + 8 * ar.s, ar.ph);
                    else mbar_try(bGamFull + 8 * gr.s, gr.ph);
                }
            }
        }
    } else if (warp >= EG_REC_WARPS) {
        // ===================== converters: raw features -> fp16 hi/lo A image =====================
        const int cw = warp - EG_REC_WARPS;
        const int q = cw & 3, h = cw >> 2;
        const int r = q * 32 + lane;
        const int nch = NCH > 0 ? NCH : nck / 2;
        const int c0 = h * nch;
        constexpr int NCMAX = NCH > 0 ? NCH : 6;
        auto lds4 = [](uint32_t a) -> float4 {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            return v;
        };
        const uint32_t sbS = smem_u32(smem + L.sb) + 16u * c0, sbB = sbS + 16u * nck;
        float4 rsc[NCMAX], rof[NCMAX];
#pragma unroll
        for (int c = 0; c < NCMAX; c++)
            if (c < nch) { rsc[c] = lds4(sbS + 16u * c); rof[c] = lds4(sbB + 16u * c); }
        const uint32_t raw_row = pin_reg(sRaw + (uint32_t)r * rw + 16u * c0);
        const uint32_t a_row = pin_reg(sA + (uint32_t)(r >> 3) * rg_stride + (uint32_t)c0 * 128u + (uint32_t)(r & 7) * 16u);
        const uint32_t ta_row = pin_reg(tmem_a + ((uint32_t)(q * 32) << 16));
        const uint32_t lane0 = pin_reg(lane == 0 ? 1u : 0u);
        Ring rr = {0, 0}, ar = {0, 0}, ir = {0, 0};
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            for (int i = 0; i < npf; i++) {
                const bool img = i > T;                               // backward frame whose statistics will be accumulated
                mbar_wait(bRawFull + 8 * rr.s, rr.ph);
                mbar_wait(bAFree + 8 * ar.s, ar.ph ^ 1u);
                if (img) mbar_wait(bImgFree + 8 * ir.s, ir.ph ^ 1u);
                tc_fence_after();
                const uint32_t src = raw_row + rr.s * frame_bytes, dst = a_row + ir.s * a_stage, ta = ta_row + ar.s * a_cols;
                float4 xr[NCMAX];                                     // all loads first: the (ordered) shared-memory loads overlap
#pragma unroll
                for (int c = 0; c < NCMAX; c++)
                    if (c < nch) xr[c] = lds4(src + 16u * c);
                if (!EXP || !(p.flags & 4)) {
#pragma unroll
                    for (int c = 0; c < NCMAX; c++) {
                        if (c < nch) {
                            const float4 x = xr[c];
                            const float2 a01 = fma2(make_float2(x.x, x.y), make_float2(rsc[c].x, rsc[c].y), make_float2(rof[c].x, rof[c].y));
                            const float2 a23 = fma2(make_float2(x.z, x.w), make_float2(rsc[c].z, rsc[c].w), make_float2(rof[c].z, rof[c].w));
                            const float2 q01 = mul2(a01, a01), q23 = mul2(a23, a23);
                            const uint32_t h0 = pack_h2(a01), h1 = pack_h2(a23), h2 = pack_h2(q01), h3 = pack_h2(q23);
                            const uint32_t l0 = pack_h2(residual_h2(a01, h0)), l1 = pack_h2(residual_h2(a23, h1));
                            const uint32_t l2 = pack_h2(residual_h2(q01, h2)), l3 = pack_h2(residual_h2(q23, h3));
                            // emission operand in TMEM: chunk cg of row r = 4 columns, K step cg / 2: [hi c even | hi c odd | lo c even | lo c odd]
                            const uint32_t cg = (uint32_t)(c0 + c);
                            const uint32_t tc = ta + 16u * (cg >> 1) + 4u * (cg & 1u);
                            tmem_st4(tc, h0, h1, h2, h3);
                            tmem_st4(tc + 8u, l0, l1, l2, l3);
                            if (img) {                                // the same halves as the shared-memory image of the statistics product
                                sts128(dst + 128u * c, h0, h1, h2, h3);
                                sts128(dst + a_plane + 128u * c, l0, l1, l2, l3);
                            }
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                if (img) fence_proxy_async();                         // generic-proxy stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane0) { mbar_arrive(bAFull + 8 * ar.s); mbar_arrive(bRawEmpty + 8 * rr.s); }
                rr.next((uint32_t)nraw);
                ar.next(EG_T_STAGES);
                if (img) ir.next(EG_I_STAGES);
            }
        }
    } else {
        // ===================== recursions: one thread per utterance =====================
        const int q = warp, r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        float *scr = p.scratch + (size_t)blockIdx.x * T * 8 * TC_ROWS + r;      // [t][8][row]
        Ring cr = {0, 0}, gr = {0, 0}, sr = {0, 0};
        const uint32_t g_row = (uint32_t)(r >> 3) * 256u + (uint32_t)(r & 7) * 16u;
        float racc[8];                                     // statistics of feature row r (this thread's TMEM lane), states 1..8
#pragma unroll
        for (int j = 0; j < 8; j++) racc[j] = 0.f;

        auto fetch = [&](float (&e)[8]) {
            mbar_wait(bAccFull + 8 * cr.s, cr.ph);
            tc_fence_after();
            uint32_t ev[16];
            tmem_ld16(tmem_acc + lane_sel + cr.s * 16u, ev);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bAccEmpty + 8 * cr.s);
            cr.next(EG_ACC);
#pragma unroll
            for (int j = 0; j < 8; j++) e[j] = __uint_as_float(ev[j]) + __uint_as_float(ev[8 + j]);
        };
        auto drain = [&]() {      // one finished statistics group: TMEM -> registers
            mbar_wait(bStFull + 8 * sr.s, sr.ph);
            tc_fence_after();
            uint32_t sv[16];
            tmem_ld16(tmem_st + lane_sel + sr.s * 16u, sv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bStEmpty + 8 * sr.s);
            sr.next(2);
#pragma unroll
            for (int j = 0; j < 8; j++) racc[j] += __uint_as_float(sv[j]) + __uint_as_float(sv[8 + j]);
        };

        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int m = p.tile_model[tile];
            const bool valid = r < p.tile_rows[tile];
            const int u = p.tile_u0[tile] + r;
            // transition constants of the tile's model, in registers
            const float4 c03 = sTr[m * TC_TRQ + 0], c47 = sTr[m * TC_TRQ + 1], cm = sTr[m * TC_TRQ + 2];
            const float4 st03 = sTr[m * TC_TRQ + 3], st47 = sTr[m * TC_TRQ + 4];
            const float4 bd03 = sTr[m * TC_TRQ + 5], bd47 = sTr[m * TC_TRQ + 6];
            const float b_ent = sTr[m * TC_TRQ + 7].x;
            const float cadv[8] = {0.f, c03.x, c03.y, c03.z, c03.w, c47.x, c47.y, c47.z};
            const float stay[8] = {st03.x, st03.y, st03.z, st03.w, st47.x, st47.y, st47.z, st47.w};
            const float badv[7] = {bd03.x, bd03.y, bd03.z, bd03.w, bd47.x, bd47.y, bd47.z};
            const float lb0 = cm.y;

            // ---------------- forward (custom_hmm.py:176-211), U_j = alpha-hat_j + ln A[j,j] ----------------
            float U[8], ax = -INFINITY, base = 0.f;
            float best_mx = 0.f, best_base = 0.f;
            double ll = 0.0;
            bool exit_ok = false, xi_live = true;
#pragma unroll
            for (int j = 0; j < 8; j++) U[j] = -INFINITY;
            for (int t = 0; t < T; t++) {
                float e[8];
                fetch(e);
                if (EXP && (p.flags & 2)) continue;                            // tuning experiment: pipeline without the recursion arithmetic
                float *sp = scr + (size_t)t * 8 * TC_ROWS;
                if (t == 0) {
                    U[0] = lb0 + e[0];
                } else {
                    const float nx = U[7] + c47.w;
#pragma unroll
                    for (int j = 7; j >= 1; j--) U[j] = eg_lae(U[j - 1] + cadv[j], U[j]) + e[j];
                    const float ent = (t == 1) ? lb0 - base : -INFINITY;
                    U[0] = eg_lae(ent, U[0]) + e[0];
                    ax = nx;
                }
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = U[j] - stay[j];
                float mx = fmaxf(fmaxf(a[0], a[1]), ax);
                mx = fmaxf(fmaxf(a[2], a[3]), mx);
                mx = fmaxf(fmaxf(a[4], a[5]), mx);
                mx = fmaxf(fmaxf(a[6], a[7]), mx);
                if (mx + base > best_mx + best_base) { best_mx = mx; best_base = base; }
                if (t == T - 1) {
                    float rr = (T == 1) ? 0.f - base : -INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; j++) rr = eg_lae(rr, a[j]);
                    rr = eg_lae(rr, ax);
                    ll = ((double)rr + (double)base) - ((double)best_mx + (double)best_base);
                    exit_ok = (T > 1) && (ax > -INFINITY);
                    xi_live = !(ax - rr < -745.13f);      // SURVEY D10 (see k_estep_fused): the reference's float64 xi underflows for the whole utterance
                }
                const float sh = (t == 0) ? 0.f : rintf(fminf(fmaxf(mx, -4194304.f), 4194304.f));
#pragma unroll
                for (int j = 0; j < 8; j++) { U[j] -= sh; a[j] -= sh; }
                ax -= sh; base += sh;
#pragma unroll
                for (int j = 0; j < 8; j++) sp[j * TC_ROWS] = a[j];
            }
            if (valid) p.loglik[u] = ll;

            // ---------------- backward + gamma + xi sums (custom_hmm.py:213-322) with the statistics product ----------------
            float gG[8], gX[8], b[8], en[8];
            float glast = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) { gG[j] = 0.f; gX[j] = 0.f; b[j] = -INFINITY; en[j] = 0.f; }
            float bx = 0.f;
            float atn[8];
#pragma unroll
            for (int j = 0; j < 8; j++) atn[j] = 0.f;
            if (T >= 2) {
                const float *sp = scr + (size_t)(T - 2) * 8 * TC_ROWS;
#pragma unroll
                for (int j = 0; j < 8; j++) atn[j] = sp[j * TC_ROWS];
            }
            int sf = 0, drained = 0;                               // statistics frames written / groups drained in this tile
            for (int t = T - 1; t >= 0; t--) {
                if (t >= EG_PF) {
                    const float *pp = scr + (size_t)(t - EG_PF) * 8 * TC_ROWS;
#pragma unroll
                    for (int j = 0; j < 8; j++) asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + j * TC_ROWS));
                }
                float e[8];
                fetch(e);
                if (t == T - 1) {
                    glast = exit_ok ? 0.f : NAN;
                } else if (EXP && (p.flags & 2)) {                           // tuning experiment: zero posteriors, barrier protocol only
                    mbar_wait(bGamFree + 8 * gr.s, gr.ph ^ 1u);
                    const uint32_t gb = sG + gr.s * 4096u + g_row;
                    sts128(gb, 0, 0, 0, 0); sts128(gb + 128u, 0, 0, 0, 0);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bGamFull + 8 * gr.s);
                    gr.next(EG_GAM);
                    sf++;
                    if (sf >= (drained + 1) * EG_GROUP + 2) { drain(); drained++; }
                } else {
                    float at[8], self[8], nb[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { at[j] = atn[j]; self[j] = en[j] + b[j]; }
                    if (t >= 1) {
                        const float *sp = scr + (size_t)(t - 1) * 8 * TC_ROWS;
#pragma unroll
                        for (int j = 0; j < 8; j++) atn[j] = sp[j * TC_ROWS];
                    }
#pragma unroll
                    for (int j = 0; j < 7; j++) nb[j] = eg_lae(self[j], badv[j] + self[j + 1]);
                    nb[7] = eg_lae(self[7], bd47.w + bx);
                    const float b0 = b_ent + self[0];
                    float lg[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) lg[j] = at[j] + nb[j];
                    const float xs7 = at[7] + self[7];
                    float mxl = fmaxf(fmaxf(lg[0], lg[1]), lg[2]);
                    mxl = fmaxf(fmaxf(lg[3], lg[4]), mxl);
                    mxl = fmaxf(fmaxf(lg[5], lg[6]), mxl);
                    mxl = fmaxf(lg[7], mxl);
                    const float ent = (t == 0) ? b0 : -INFINITY;
                    mxl = fmaxf(mxl, ent);
                    float pj[8], sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; j++) { pj[j] = eg_exp(lg[j] - mxl); sum += pj[j]; }
                    const float pe = eg_exp(ent - mxl);
                    const float q7 = eg_exp(xs7 - mxl);
                    const float xsum = (sum - pj[7]) + q7 + pe;
                    sum += pe;
                    const float inv = 1.0f / sum;
                    const float xinv = (xi_live && mxl > -INFINITY && xsum > 0.f) ? 1.0f / xsum : 0.f;
                    float gm[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        gm[j] = pj[j] * inv;
                        gG[j] += gm[j];
                        const float xq = (mxl > -INFINITY) ? eg_exp((at[j] + self[j]) - mxl) : 0.f;
                        gX[j] += xq * xinv;
                    }
                    // posteriors of frame t -> fp16 hi/lo B operand of the statistics product (rows past the model's last
                    // utterance contribute exact zeros)
                    {
                        const float vz = valid ? 1.f : 0.f;
                        const float2 g01 = make_float2(gm[0] * vz, gm[1] * vz), g23 = make_float2(gm[2] * vz, gm[3] * vz);
                        const float2 g45 = make_float2(gm[4] * vz, gm[5] * vz), g67 = make_float2(gm[6] * vz, gm[7] * vz);
                        const uint32_t h0 = pack_h2(g01), h1 = pack_h2(g23), h2 = pack_h2(g45), h3 = pack_h2(g67);
                        mbar_wait(bGamFree + 8 * gr.s, gr.ph ^ 1u);
                        const uint32_t gb = sG + gr.s * 4096u + g_row;
                        sts128(gb, h0, h1, h2, h3);
                        sts128(gb + 128u, pack_h2(residual_h2(g01, h0)), pack_h2(residual_h2(g23, h1)), pack_h2(residual_h2(g45, h2)),
                               pack_h2(residual_h2(g67, h3)));
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bGamFull + 8 * gr.s);
                        gr.next(EG_GAM);
                        sf++;
                    }
                    float mb = fmaxf(fmaxf(nb[0], nb[1]), nb[2]);
                    mb = fmaxf(fmaxf(nb[3], nb[4]), mb);
                    mb = fmaxf(fmaxf(nb[5], nb[6]), mb);
                    mb = fmaxf(nb[7], mb);
                    const bool fin = mb > -INFINITY && mb < INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; j++) b[j] = fin ? nb[j] - mb : nb[j];
                    bx = -INFINITY;
                    // a group finished two frames ago has certainly been multiplied by now: drain it without waiting
                    if (sf >= (drained + 1) * EG_GROUP + 2) { drain(); drained++; }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) en[j] = e[j];
            }
            const int ngroups = (T - 1 + EG_GROUP - 1) / EG_GROUP;
            for (; drained < ngroups; drained++) drain();
            {   // per-tile partial sums of this thread's feature row
                float4 *tp = reinterpret_cast<float4 *>(p.tpart + ((size_t)tile * TC_ROWS + r) * 8);
                tp[0] = make_float4(racc[0], racc[1], racc[2], racc[3]);
                tp[1] = make_float4(racc[4], racc[5], racc[6], racc[7]);
#pragma unroll
                for (int j = 0; j < 8; j++) racc[j] = 0.f;
            }
            if (valid) {
                float *us = p.ustats + (size_t)u * 24;
#pragma unroll
                for (int j = 0; j < 8; j++) { us[j] = gG[j]; us[8 + j] = gX[j]; us[16 + j] = gG[j] + glast; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EG_MMA_WARP) tmem_dealloc(tmem_base, tcols);
}

// ------------------------------------------------------------------------------------------------
// per-utterance (G, Xi, occ) triples -> stats (fixed order): grid = (24, M), utterances of model m are contiguous
__global__ void k_eg_reduce_triples(const int32_t *__restrict__ model_start, int S, const float *__restrict__ ustats,
                                    double *__restrict__ stats, int64_t stride) {
    const int m = blockIdx.y, qd = blockIdx.x;
    __shared__ double s_acc[256];
    double s = 0.0;
    for (int u = model_start[m] + threadIdx.x; u < model_start[m + 1]; u += blockDim.x) s += (double)ustats[(size_t)u * 24 + qd];
    s_acc[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) s_acc[threadIdx.x] += s_acc[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) stats[(size_t)m * stride + (size_t)(qd / 8) * S + (qd % 8) + 1] = s_acc[0];
}

// per-tile sums of the standardised features -> statistics around the states' current means (float64, fixed order).
// Feature row L of the A image: chunk c = L / 8, e = L % 8: e < 4 -> x'_{4c+e}, e >= 4 -> x'^2_{4c+e-4}; dim index D is the constant
// slot (x' = 1), whose first-order row carries sum gamma.  x' = x * sf + bf exactly as the kernel standardises.
__global__ void k_eg_reduce_features(const int32_t *__restrict__ tile_first, int M, int D, int S, const float *__restrict__ tpart,
                                     const float *__restrict__ sb, int nck, const double *__restrict__ mean, double *__restrict__ stats,
                                     int64_t stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * 8 * D) return;
    const int m = idx / (8 * D), j = (idx / D) % 8, d = idx % D;
    const int L1 = (d / 4) * 8 + (d % 4), L2 = L1 + 4, LG = (D / 4) * 8 + (D % 4);
    double t1 = 0.0, t2 = 0.0, g = 0.0;
    for (int tile = tile_first[m]; tile < tile_first[m + 1]; tile++) {
        const float *tp = tpart + (size_t)tile * TC_ROWS * 8;
        t1 += (double)tp[L1 * 8 + j]; t2 += (double)tp[L2 * 8 + j]; g += (double)tp[LG * 8 + j];
    }
    const double sf = (double)sb[d], bf = (double)sb[4 * nck + d];
    const double mu = mean[((size_t)m * S + j + 1) * D + d];
    // sum g x = (t1 - bf g) / sf ; sum g x^2 = (t2 - 2 bf t1 + bf^2 g) / sf^2
    const double sx = (t1 - bf * g) / sf, sxx = (t2 - 2.0 * bf * t1 + bf * bf * g) / (sf * sf);
    double *st = stats + (size_t)m * stride + 3 * S;
    st[(size_t)(j + 1) * D + d] = sx - mu * g;
    st[(size_t)S * D + (size_t)(j + 1) * D + d] = sxx - 2.0 * mu * sx + mu * mu * g;
}

// ------------------------------------------------------------------------------------------------
extern "C" int sapr_estep_grouped(sapr_ctx *ctx, sapr_models *m, const float *X, int ldx, int T, int B,
                                  const int32_t *model_start_host, double *stats, double *loglik) {
    if (!ctx || !m || !X || !model_start_host || !stats || !loglik) return SAPR_E_INVALID;
    if (!m->valid) SAPR_FAIL(ctx, SAPR_E_INVALID, "estep_grouped: model parameters not set");
    if (!m->tc_image || !sapr_tc_eligible(m))
        SAPR_FAIL(ctx, SAPR_E_RANGE, "estep_grouped: needs DIAG emission, ENTRY_EXIT topology, N = 8, M <= 12, D <= 47");
    if (ldx % 4 || ldx < m->Dp || ((uintptr_t)X & 15u)) SAPR_FAIL(ctx, SAPR_E_INVALID, "estep_grouped: X rows must be 16-byte aligned, ldx >= D padded");
    if (T < 16) SAPR_FAIL(ctx, SAPR_E_RANGE, "estep_grouped: T >= 16 (shorter utterances: sapr_estep)");
    const int M = m->M, D = m->D, S = m->S;
    if (model_start_host[0] != 0 || model_start_host[M] != B) SAPR_FAIL(ctx, SAPR_E_INVALID, "estep_grouped: model_start must span [0, B]");
    const int64_t stride = sapr_stats_stride(m->N, D);
    SAPR_CUDA(ctx, cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)M * stride, ctx->stream));
    if (B <= 0) return SAPR_OK;
    int nck, ncols;
    const size_t img_bytes = sapr_tc_image_bytes(m, &nck, &ncols);
    (void)img_bytes;
    auto al256 = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t w = al256((size_t)2 * (ncols / 8) * nck * 64 * sizeof(__half));
    const size_t g = al256((size_t)8 * nck * sizeof(float));
    // tiles: 128 consecutive utterances of one model
    std::vector<int32_t> tab;
    std::vector<int32_t> tmodel, tu0, trows, tfirst(M + 1, 0);
    for (int mi = 0; mi < M; mi++) {
        const int a = model_start_host[mi], b = model_start_host[mi + 1];
        if (b < a) SAPR_FAIL(ctx, SAPR_E_INVALID, "estep_grouped: model_start must be non-decreasing");
        tfirst[mi] = (int32_t)tmodel.size();
        for (int u = a; u < b; u += TC_ROWS) { tmodel.push_back(mi); tu0.push_back(u); trows.push_back(std::min(TC_ROWS, b - u)); }
    }
    tfirst[M] = (int32_t)tmodel.size();
    const int ntiles = (int)tmodel.size();
    tab.reserve(3 * ntiles + 2 * (M + 1));
    tab.insert(tab.end(), tmodel.begin(), tmodel.end());
    tab.insert(tab.end(), tu0.begin(), tu0.end());
    tab.insert(tab.end(), trows.begin(), trows.end());
    tab.insert(tab.end(), tfirst.begin(), tfirst.end());
    tab.insert(tab.end(), model_start_host, model_start_host + M + 1);
    int rc;
    // the table lives in its own workspace slot and is uploaded only when the grouping changed: repeated iterations
    // over the same batch enqueue without touching the host
    if (tab != ctx->eg_tab || !ctx->ws[6] || ctx->ws_bytes[6] < tab.size() * sizeof(int32_t)) {
        if ((rc = sapr_ws_reserve(ctx, 6, tab.size() * sizeof(int32_t)))) return rc;
        if ((rc = sapr_pin_reserve(ctx, 2, tab.size() * sizeof(int32_t)))) return rc;
        SAPR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // the previous table may still be in use
        memcpy(ctx->pin[2], tab.data(), tab.size() * sizeof(int32_t));
        SAPR_CUDA(ctx, cudaMemcpyAsync(ctx->ws[6], ctx->pin[2], tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        ctx->eg_tab = tab;
    }
    const int32_t *d_tab = (const int32_t *)ctx->ws[6];
    const int grid = std::min(ntiles, ctx->sm_count);
    if ((rc = sapr_ws_reserve(ctx, 7, (size_t)grid * T * 8 * TC_ROWS * sizeof(float)))) return rc;
    if ((rc = sapr_ws_reserve(ctx, 1, sizeof(float) * (size_t)B * 24))) return rc;
    if ((rc = sapr_ws_reserve(ctx, 4, sizeof(float) * (size_t)ntiles * TC_ROWS * 8))) return rc;
    const uint32_t rw = (uint32_t)(nck + 1) * 16u;
    int nraw = EG_MAX_RAW;
    while (nraw > 2 && eg_smem_layout(M, nck, nraw, rw).total > 227 * 1024) nraw--;
    const EgSmem L = eg_smem_layout(M, nck, nraw, rw);
    if (L.total > 227 * 1024) SAPR_FAIL(ctx, SAPR_E_RANGE, "estep_grouped: feature rows too wide for shared memory");
    CUtensorMap tmap;
    {
        const uint64_t dim[3] = {(uint64_t)ldx, (uint64_t)T, (uint64_t)B};
        const uint64_t str[2] = {(uint64_t)ldx * 4u, (uint64_t)T * ldx * 4u};
        const uint32_t box[3] = {rw / 4u, 1u, (uint32_t)TC_ROWS};
        if ((rc = sapr_tmap_f32_3d(ctx, &tmap, X, dim, str, box))) return rc;
    }
    EgParams prm;
    prm.T = T; prm.B = B; prm.M = M; prm.D = D; prm.nck = nck; prm.ntiles = ntiles; prm.nraw = nraw;
    prm.tile_model = d_tab; prm.tile_u0 = d_tab + ntiles; prm.tile_rows = d_tab + 2 * ntiles;
    prm.wimg = (const __half *)m->tc_image; prm.w_plane_halves = (uint32_t)((ncols / 8) * nck * 64);
    prm.sb = (const float *)((const char *)m->tc_image + w);
    prm.trp = (const float4 *)((const char *)m->tc_image + w + g);
    prm.scratch = (float *)ctx->ws[7]; prm.ustats = (float *)ctx->ws[1]; prm.loglik = loglik; prm.tpart = (float *)ctx->ws[4];
    prm.rw = rw; prm.trace = nullptr;
    prm.flags = getenv("SAPR_EG_EXP") ? atoi(getenv("SAPR_EG_EXP")) : 0;
    prm.pfd = getenv("SAPR_EG_PFD") ? atoi(getenv("SAPR_EG_PFD")) : 0;
    auto kern = (nck == 10) ? (prm.flags ? k_estep_grouped<5, true> : k_estep_grouped<5>) : (nck == 4) ? k_estep_grouped<2> : k_estep_grouped<0>;
    SAPR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    {
        ProfScope ps(ctx, 2);
        kern<<<grid, EG_THREADS, L.total, ctx->stream>>>(prm, tmap);
    }
    SAPR_LAUNCH_CHECK(ctx);
    const int32_t *d_tfirst = d_tab + 3 * ntiles, *d_mstart = d_tfirst + (M + 1);
    {
        ProfScope ps(ctx, 3);
        k_eg_reduce_triples<<<dim3(24, M), 256, 0, ctx->stream>>>(d_mstart, S, (const float *)ctx->ws[1], stats, stride);
    }
    SAPR_LAUNCH_CHECK(ctx);
    k_eg_reduce_features<<<(M * 8 * D + 127) / 128, 128, 0, ctx->stream>>>(d_tfirst, M, D, S, (const float *)ctx->ws[4], prm.sb, nck, m->mean,
                                                                          stats, stride);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}
