// mfcc.cu -- MFCC front-end as fused kernels (secondary path).
//
// Replaces assignment2/mfcc_extract.py:10-27, i.e. librosa.feature.mfcc(y, sr, n_mfcc=13,
// win_length, hop_length, window="hamming", center=True): framing with centre zero-padding ->
// periodic Hamming window zero-padded to n_fft -> |rFFT|^2 -> mel filterbank -> power_to_db (with
// the utterance-global top_db clamp) -> orthonormal DCT-II -> first n_mfcc rows.  librosa is not
// vendored in the reference; the semantics follow SURVEY.md Appendix C and are parameterised so that
// BASELINE cfg 5 (16 kHz, 400/160, 512-pt, 26 HTK mels, pre-emphasis 0.97, natural log) is the
// same kernel.
//
// k_mfcc_logmel: persistent CTAs, one WARP per pair of frames: samples (+pre-emphasis, window) of the two real
//                frames form the real / imaginary part of ONE complex FFT (three radix-2 stages per register pass
//                over a padded shared-memory buffer, __syncwarp only), split spectra, power, sparse triangular mel
//                filters, log -> logmel[frame][n_mels] and the per-utterance maximum (ordered-int atomicMax).
// k_mfcc_dct:    clamp to max - top_db, DCT-II, write feats[frame][ld_out].
#include "common.cuh"

#define MFCC_PI 3.14159265358979323846
#define MFCC_PW_REG 9        /* power-spectrum values per lane held in registers when n_fft <= 512 */

extern "C" int64_t sapr_mfcc_num_frames(const sapr_mfcc_params *p, int64_t n_samples) {
    if (!p || p->hop_length <= 0) return 0;
    if (p->center) return 1 + n_samples / p->hop_length;
    if (n_samples < p->n_fft) return 0;
    return 1 + (n_samples - p->n_fft) / p->hop_length;
}

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return (i >= 0) ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7fffffff); }

__global__ void k_mfcc_init_max(int *umax, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) umax[i] = float_to_ordered(-INFINITY);
}

// frame -> utterance by binary search over the frame offsets (replaces a host-built table of one int per frame)
__global__ void k_mfcc_utt_of_frame(const int64_t *__restrict__ frame_off, int B, int64_t total_frames, int32_t *__restrict__ uof) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= total_frames) return;
    int lo = 0, hi = B;                       // largest u with frame_off[u] <= f
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (frame_off[mid] <= f) lo = mid; else hi = mid;
    }
    uof[f] = lo;
}

// One WARP per PAIR of frames (blockDim.x / 32 pairs per CTA).  Frame a goes into the real part and frame b into the
// imaginary part of ONE complex transform (both spectra fall out of Z[k] and conj(Z[N-k])), the radix-2 stages are taken
// three at a time in registers (same butterflies, a third of the shared-memory passes), and only __syncwarp separates the passes.
__global__ void __launch_bounds__(256)
k_mfcc_logmel(const float *__restrict__ audio, const int64_t *__restrict__ sample_off,
              const int64_t *__restrict__ frame_off, const int32_t *__restrict__ utt_of_frame, int64_t total_frames,
              int n_fft, int log2n, int hop, int center, float preemph,
              const float *__restrict__ g_window, const float2 *__restrict__ g_twiddle, int n_mels,
              const int32_t *__restrict__ g_mel_lo, const int32_t *__restrict__ g_mel_hi,
              const int32_t *__restrict__ g_mel_ptr, const float *__restrict__ g_mel_w, int n_mel_w, int log_db,
              float *__restrict__ logmel, int *__restrict__ umax) {
    extern __shared__ float2 s_all[];   // tables, then per warp: n_fft complex (one pad slot per 16: index i -> i + i / 16) and 2 x (n_fft/2 + 1) power values
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const int nb = n_fft / 2 + 1;
    // persistent CTA: the parameter tables are staged in shared memory once, then the warps loop over frame pairs
    float2 *twiddle = s_all;                                             // n_fft / 2
    float *window = reinterpret_cast<float *>(twiddle + n_fft / 2);      // n_fft
    float *mel_w = window + n_fft;                                       // n_mel_w
    int *mel_lo = reinterpret_cast<int *>(mel_w + n_mel_w), *mel_hi = mel_lo + n_mels, *mel_ptr = mel_hi + n_mels;
    const size_t tab_f2 = ((size_t)n_fft + n_fft + n_mel_w + 3 * n_mels + 1) / 2 + 1;      // table size in float2 units (8-byte aligned)
    for (int i = threadIdx.x; i < n_fft / 2; i += blockDim.x) twiddle[i] = g_twiddle[i];
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) window[i] = g_window[i];
    for (int i = threadIdx.x; i < n_mel_w; i += blockDim.x) mel_w[i] = g_mel_w[i];
    for (int i = threadIdx.x; i < n_mels; i += blockDim.x) { mel_lo[i] = g_mel_lo[i]; mel_hi[i] = g_mel_hi[i]; mel_ptr[i] = g_mel_ptr[i]; }
    __syncthreads();
    const bool pw_alias = nb <= 32 * MFCC_PW_REG;                                            // power values overwrite the transform buffer
    const size_t per_warp = (size_t)n_fft + (size_t)(n_fft / 16) + (pw_alias ? 0 : (size_t)((2 * nb + 1) / 2));   // in float2 units
    float2 *s_x = s_all + tab_f2 + (size_t)warp * per_warp;
    float *pw = reinterpret_cast<float *>(pw_alias ? s_x : s_x + n_fft + n_fft / 16);          // pw[which * nb + k]
    auto P = [](int i) { return i + (i >> 4); };      // padded index: the strided passes and the bit-reversed scatter hit distinct banks
    auto cmul = [](float2 b, float2 w) { return make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x); };
    for (int64_t fa = ((int64_t)blockIdx.x * wpc + warp) * 2; fa < total_frames; fa += (int64_t)gridDim.x * wpc * 2) {
    const bool two = fa + 1 < total_frames;
    int uu[2], len[2], start[2];
    const float *ap[2];                              // utterance base pointers; sample indices inside an utterance fit 32 bits
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const int64_t f = (w == 0 || two) ? fa + w : fa;
        uu[w] = utt_of_frame[f];
        const int64_t s0 = sample_off[uu[w]];
        ap[w] = audio + s0;
        len[w] = (w == 0 || two) ? (int)(sample_off[uu[w] + 1] - s0) : 0;
        start[w] = (int)(f - frame_off[uu[w]]) * hop - (center ? n_fft / 2 : 0);
    }
    // four consecutive samples per lane and step: one 16-byte load per frame when the chunk is aligned and inside the
    // utterance (the common case), element-wise guarded loads at the utterance boundaries; all chunks of a lane are in flight
    // together (n_fft / 128 steps)
    bool al16[2];
#pragma unroll
    for (int w = 0; w < 2; w++) al16[w] = ((reinterpret_cast<uintptr_t>(ap[w] + start[w]) & 15) == 0);
#pragma unroll 4
    for (int c = lane; c < n_fft / 4; c += 32) {
        const int i = 4 * c;
        const float4 wv = *reinterpret_cast<const float4 *>(window + i);
        const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
        float y[2][4];
        const bool any = (wv.x != 0.0f) || (wv.y != 0.0f) || (wv.z != 0.0f) || (wv.w != 0.0f);
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const int sidx = start[w] + i;
            float x[4], xp;
            if (!any || len[w] == 0) {
                x[0] = x[1] = x[2] = x[3] = 0.0f; xp = 0.0f;
            } else if (al16[w] && sidx >= 1 && sidx + 3 < len[w]) {
                const float4 v = *reinterpret_cast<const float4 *>(ap[w] + sidx);
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                xp = ap[w][sidx - 1];
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) { const int sx = sidx + e; x[e] = (sx >= 0 && sx < len[w]) ? ap[w][sx] : 0.0f; }
                xp = (sidx >= 1 && sidx - 1 < len[w]) ? ap[w][sidx - 1] : 0.0f;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int sx = sidx + e;
                const float prev = (e == 0) ? xp : x[e - 1];
                const float v = (sx >= 0 && sx < len[w]) ? fmaf(-preemph, (sx > 0) ? prev : 0.0f, x[e]) : 0.0f;
                y[w][e] = v * wq[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int r = __brev((unsigned)(i + e)) >> (32 - log2n);
            s_x[r + (r >> 4)] = make_float2(y[0][e], y[1][e]);
        }
    }
    __syncwarp();
    int st = 1;
    for (; st + 2 <= log2n; st += 3) {          // stages st .. st + 2 on the 8 elements p + q h (q = 0..7) of a block of 8h, in registers
        const int h = 1 << (st - 1);
        for (int k = lane; k < n_fft / 8; k += 32) {
            const int pos = k & (h - 1);
            const int i0 = ((k - pos) << 3) + pos;
            float2 a[8];
            int j[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { j[q] = P(i0 + q * h); a[q] = s_x[j[q]]; }
            {   // stage st: pairs (0,1) (2,3) (4,5) (6,7), one twiddle
                const float2 w = twiddle[pos * (n_fft >> st)];
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const float2 t = cmul(a[q + 1], w), x = a[q];
                    a[q] = make_float2(x.x + t.x, x.y + t.y); a[q + 1] = make_float2(x.x - t.x, x.y - t.y);
                }
            }
            {   // stage st + 1: pairs (0,2) (1,3) (4,6) (5,7), twiddles at pos and pos + h
                const float2 w0 = twiddle[pos * (n_fft >> (st + 1))], w1 = twiddle[(pos + h) * (n_fft >> (st + 1))];
#pragma unroll
                for (int q = 0; q < 8; q += 4) {
                    float2 t = cmul(a[q + 2], w0), x = a[q];
                    a[q] = make_float2(x.x + t.x, x.y + t.y); a[q + 2] = make_float2(x.x - t.x, x.y - t.y);
                    t = cmul(a[q + 3], w1); x = a[q + 1];
                    a[q + 1] = make_float2(x.x + t.x, x.y + t.y); a[q + 3] = make_float2(x.x - t.x, x.y - t.y);
                }
            }
            {   // stage st + 2: pairs (q, q + 4), twiddles at pos + q h
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float2 t = cmul(a[q + 4], twiddle[(pos + q * h) * (n_fft >> (st + 2))]), x = a[q];
                    a[q] = make_float2(x.x + t.x, x.y + t.y); a[q + 4] = make_float2(x.x - t.x, x.y - t.y);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) s_x[j[q]] = a[q];
        }
        __syncwarp();
    }
    for (; st + 1 <= log2n; st += 2) {          // stages st and st + 1 on the 4 elements p, p+h, p+2h, p+3h of a block of 4h
        const int h = 1 << (st - 1);
        for (int k = lane; k < n_fft / 4; k += 32) {
            const int pos = k & (h - 1);
            const int i0 = ((k - pos) << 2) + pos;
            const float2 w1 = twiddle[pos * (n_fft >> st)];
            const float2 w2 = twiddle[pos * (n_fft >> (st + 1))], w3 = twiddle[(pos + h) * (n_fft >> (st + 1))];
            const int j0 = P(i0), j1 = P(i0 + h), j2 = P(i0 + 2 * h), j3 = P(i0 + 3 * h);
            const float2 a0 = s_x[j0], a1 = s_x[j1], a2 = s_x[j2], a3 = s_x[j3];
            float2 t = cmul(a1, w1);
            const float2 b0 = make_float2(a0.x + t.x, a0.y + t.y), b1 = make_float2(a0.x - t.x, a0.y - t.y);
            t = cmul(a3, w1);
            const float2 b2 = make_float2(a2.x + t.x, a2.y + t.y), b3 = make_float2(a2.x - t.x, a2.y - t.y);
            t = cmul(b2, w2);
            s_x[j0] = make_float2(b0.x + t.x, b0.y + t.y);
            s_x[j2] = make_float2(b0.x - t.x, b0.y - t.y);
            t = cmul(b3, w3);
            s_x[j1] = make_float2(b1.x + t.x, b1.y + t.y);
            s_x[j3] = make_float2(b1.x - t.x, b1.y - t.y);
        }
        __syncwarp();
    }
    if (st <= log2n) {                           // odd number of stages: one plain radix-2 stage is left
        const int half = 1 << (st - 1), tws = n_fft >> st;
        for (int k = lane; k < n_fft / 2; k += 32) {
            const int pos = k & (half - 1);
            const int i0 = P(((k - pos) << 1) + pos), i1 = P(((k - pos) << 1) + pos + half);
            const float2 t = cmul(s_x[i1], twiddle[pos * tws]);
            const float2 a = s_x[i0];
            s_x[i0] = make_float2(a.x + t.x, a.y + t.y);
            s_x[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncwarp();
    }
    // split the two real spectra: A[k] = (Z[k] + conj(Z[N-k])) / 2, B[k] = (Z[k] - conj(Z[N-k])) / (2i).  For n_fft <= 512 the
    // power values go through registers and overwrite the transform buffer (2 KB less shared memory per warp = more warps per SM)
    if (pw_alias) {
        float pa[MFCC_PW_REG], pb[MFCC_PW_REG];
#pragma unroll
        for (int q = 0; q < MFCC_PW_REG; q++) {
            const int k = lane + 32 * q;
            pa[q] = 0.0f; pb[q] = 0.0f;
            if (k < nb) {
                const float2 z = s_x[P(k)], zn = s_x[P((n_fft - k) & (n_fft - 1))];
                const float ar = 0.5f * (z.x + zn.x), ai = 0.5f * (z.y - zn.y);
                const float br = 0.5f * (z.y + zn.y), bi = 0.5f * (zn.x - z.x);
                pa[q] = ar * ar + ai * ai; pb[q] = br * br + bi * bi;
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < MFCC_PW_REG; q++) {
            const int k = lane + 32 * q;
            if (k < nb) { pw[k] = pa[q]; pw[nb + k] = pb[q]; }
        }
    } else {
        for (int k = lane; k < nb; k += 32) {
            const float2 z = s_x[P(k)], zn = s_x[P((n_fft - k) & (n_fft - 1))];
            const float ar = 0.5f * (z.x + zn.x), ai = 0.5f * (z.y - zn.y);
            const float br = 0.5f * (z.y + zn.y), bi = 0.5f * (zn.x - z.x);
            pw[k] = ar * ar + ai * ai;
            pw[nb + k] = br * br + bi * bi;
        }
    }
    __syncwarp();
    for (int e = lane; e < 2 * n_mels; e += 32) {
        const int w = e / n_mels, mth = e - w * n_mels;
        if (w == 1 && !two) break;
        float acc = 0.0f;
        const int lo = mel_lo[mth];
        const float *wt = mel_w + mel_ptr[mth];
        const float *pp = pw + w * nb;
        for (int k = lo; k < mel_hi[mth]; k++) acc += wt[k - lo] * pp[k];
        acc = fmaxf(acc, 1e-10f);
        const float lv = log_db ? 10.0f * log10f(acc) : logf(acc);
        logmel[(fa + w) * n_mels + mth] = lv;
        if (log_db) atomicMax(umax + uu[w], float_to_ordered(lv));       // only the top_db clamp reads it
    }
    __syncwarp();                                 // the next pair reuses this warp's buffers
    }
}

__global__ void k_mfcc_dct(const float *__restrict__ logmel, const int32_t *__restrict__ utt_of_frame,
                           const int *__restrict__ umax, int n_mels, int n_mfcc, float top_db,
                           const float *__restrict__ dct, float *__restrict__ feats, int ld_out, int64_t total_frames) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_frames * n_mfcc) return;
    const int64_t f = idx / n_mfcc;
    const int k = (int)(idx % n_mfcc);
    const float lo = (top_db > 0.0f) ? ordered_to_float(umax[utt_of_frame[f]]) - top_db : -INFINITY;
    const float *lm = logmel + f * n_mels;
    const float *dk = dct + (size_t)k * n_mels;
    float acc = 0.0f;
    for (int m = 0; m < n_mels; m++) acc += dk[m] * fmaxf(lm[m], lo);
    feats[f * ld_out + k] = acc;
}

static double hz_to_mel(double f, int slaney) {
    if (!slaney) return 2595.0 * log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return (f >= min_log_hz) ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m, int slaney) {
    if (!slaney) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return (m >= min_log_mel) ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

extern "C" int sapr_mfcc(sapr_ctx *ctx, const sapr_mfcc_params *p, const float *audio, const int64_t *sample_offsets_host,
                         int B, float *feats, int ld_out, int64_t *feat_offsets_host) {
    if (!ctx || !p || !audio || !sample_offsets_host || !feats || !feat_offsets_host || B <= 0) return SAPR_E_INVALID;
    int log2n = 0;
    while ((1 << log2n) < p->n_fft) log2n++;
    if ((1 << log2n) != p->n_fft || p->n_fft < 32 || p->n_fft > 4096) SAPR_FAIL(ctx, SAPR_E_RANGE, "mfcc: n_fft must be a power of two in [32, 4096]");
    if (p->win_length <= 0 || p->win_length > p->n_fft || p->hop_length <= 0) SAPR_FAIL(ctx, SAPR_E_INVALID, "mfcc: bad window/hop");
    if (p->n_mels <= 0 || p->n_mfcc <= 0 || p->n_mfcc > p->n_mels || ld_out < p->n_mfcc) SAPR_FAIL(ctx, SAPR_E_INVALID, "mfcc: bad n_mels/n_mfcc/ld_out");
    const int n_fft = p->n_fft, nb = n_fft / 2 + 1, n_mels = p->n_mels, n_mfcc = p->n_mfcc;
    // ---- host tables (float64 -> float32) ----
    std::vector<float> window(n_fft, 0.0f);
    {   // periodic Hamming (scipy get_window(..., fftbins=True)), centred zero-padding to n_fft (librosa pad_center)
        const int L = p->win_length, lpad = (n_fft - L) / 2;
        for (int i = 0; i < L; i++) window[lpad + i] = (float)(0.54 - 0.46 * cos(2.0 * MFCC_PI * i / L));
    }
    std::vector<float2> tw(n_fft / 2);
    for (int k = 0; k < n_fft / 2; k++) {
        const double a = -2.0 * MFCC_PI * k / n_fft;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // mel filterbank (librosa.filters.mel): triangles between n_mels+2 mel-spaced points
    const double fmax = (p->fmax > 0) ? p->fmax : p->sample_rate / 2.0, fmin = p->fmin;
    std::vector<double> pts(n_mels + 2);
    const double m_lo = hz_to_mel(fmin, p->mel_slaney), m_hi = hz_to_mel(fmax, p->mel_slaney);
    for (int i = 0; i < n_mels + 2; i++) pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1), p->mel_slaney);
    std::vector<int32_t> lo(n_mels), hi(n_mels), ptr(n_mels);
    std::vector<float> w;
    for (int mth = 0; mth < n_mels; mth++) {
        int first = -1, last = -1;
        std::vector<float> row(nb, 0.0f);
        for (int k = 0; k < nb; k++) {
            const double fk = (double)k * p->sample_rate / n_fft;
            const double lower = (fk - pts[mth]) / (pts[mth + 1] - pts[mth]);
            const double upper = (pts[mth + 2] - fk) / (pts[mth + 2] - pts[mth + 1]);
            double v = std::max(0.0, std::min(lower, upper));
            if (p->mel_slaney) v *= 2.0 / (pts[mth + 2] - pts[mth]);
            row[k] = (float)v;
            if (v > 0) { if (first < 0) first = k; last = k; }
        }
        if (first < 0) { first = 0; last = -1; }
        lo[mth] = first; hi[mth] = last + 1; ptr[mth] = (int32_t)w.size();
        for (int k = first; k <= last; k++) w.push_back(row[k]);
    }
    if (w.empty()) w.push_back(0.0f);
    std::vector<float> dct((size_t)n_mfcc * n_mels);   // scipy.fft.dct(type=2, norm="ortho")
    for (int k = 0; k < n_mfcc; k++)
        for (int n = 0; n < n_mels; n++) {
            const double sc = (k == 0) ? sqrt(1.0 / n_mels) : sqrt(2.0 / n_mels);
            dct[(size_t)k * n_mels + n] = (float)(sc * cos(MFCC_PI * k * (2.0 * n + 1.0) / (2.0 * n_mels)));
        }
    // ---- frame bookkeeping ----
    std::vector<int64_t> foff(B + 1, 0);
    for (int u = 0; u < B; u++) foff[u + 1] = foff[u] + sapr_mfcc_num_frames(p, sample_offsets_host[u + 1] - sample_offsets_host[u]);
    const int64_t total_frames = foff[B];
    memcpy(feat_offsets_host, foff.data(), sizeof(int64_t) * (B + 1));
    if (total_frames <= 0) return SAPR_OK;
    // ---- device tables: one workspace blob (parameter tables, per-call offsets, frame -> utterance, log-mel) ----
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o_win = 0, o_tw = o_win + al(sizeof(float) * n_fft), o_lo = o_tw + al(sizeof(float2) * (n_fft / 2));
    size_t o_hi = o_lo + al(sizeof(int32_t) * n_mels), o_ptr = o_hi + al(sizeof(int32_t) * n_mels);
    size_t o_w = o_ptr + al(sizeof(int32_t) * n_mels), o_dct = o_w + al(sizeof(float) * w.size());
    size_t o_soff = o_dct + al(sizeof(float) * dct.size()), o_foff = o_soff + al(sizeof(int64_t) * (B + 1));
    size_t o_uof = o_foff + al(sizeof(int64_t) * (B + 1)), o_umax = o_uof + al(sizeof(int32_t) * total_frames);
    size_t o_lm = o_umax + al(sizeof(int) * B), total = o_lm + al(sizeof(float) * (size_t)total_frames * n_mels);
    int rc = sapr_ws_reserve(ctx, 7, total);
    if (rc) return rc;
    char *base = (char *)ctx->ws[7];
    // pageable sources: cudaMemcpyAsync returns once they are staged, so the vectors may go out of scope afterwards
#define UP(off, vec, T) SAPR_CUDA(ctx, cudaMemcpyAsync(base + off, vec.data(), sizeof(T) * vec.size(), cudaMemcpyHostToDevice, ctx->stream))
    UP(o_win, window, float); UP(o_tw, tw, float2); UP(o_lo, lo, int32_t); UP(o_hi, hi, int32_t); UP(o_ptr, ptr, int32_t);
    UP(o_w, w, float); UP(o_dct, dct, float); UP(o_foff, foff, int64_t);
#undef UP
    SAPR_CUDA(ctx, cudaMemcpyAsync(base + o_soff, sample_offsets_host, sizeof(int64_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream));
    int *umax = (int *)(base + o_umax);
    float *logmel = (float *)(base + o_lm);
    int32_t *uof = (int32_t *)(base + o_uof);
    k_mfcc_init_max<<<(B + 127) / 128, 128, 0, ctx->stream>>>(umax, B);
    SAPR_LAUNCH_CHECK(ctx);
    k_mfcc_utt_of_frame<<<(unsigned)((total_frames + 255) / 256), 256, 0, ctx->stream>>>((const int64_t *)(base + o_foff), B, total_frames, uof);
    SAPR_LAUNCH_CHECK(ctx);
    int wpc = 8;                                           // frame pairs (warps) per CTA: as many as 96 KB of shared memory hold
    const size_t per_warp = sizeof(float2) * ((size_t)n_fft + n_fft / 16 + (nb <= 32 * MFCC_PW_REG ? 0 : (2 * nb + 1) / 2));
    const size_t tab = sizeof(float2) * (((size_t)n_fft + n_fft + w.size() + 3 * n_mels + 1) / 2 + 1);
    while (wpc > 1 && tab + per_warp * wpc > 100 * 1024) wpc >>= 1;
    const size_t smem = tab + per_warp * wpc;
    SAPR_CUDA(ctx, cudaFuncSetAttribute(k_mfcc_logmel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int cta_per_sm = 1;                                    // resident CTAs (registers and shared memory): the grid is one wave
    SAPR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cta_per_sm, k_mfcc_logmel, 32 * wpc, smem));
    cta_per_sm = std::max(cta_per_sm, 1);
    k_mfcc_logmel<<<(unsigned)std::min<int64_t>((total_frames + 2 * wpc - 1) / (2 * wpc), (int64_t)ctx->sm_count * cta_per_sm), 32 * wpc, smem, ctx->stream>>>(
        audio, (const int64_t *)(base + o_soff), (const int64_t *)(base + o_foff), uof, total_frames, n_fft,
        log2n, p->hop_length, p->center, p->preemph, (const float *)(base + o_win), (const float2 *)(base + o_tw), n_mels,
        (const int32_t *)(base + o_lo), (const int32_t *)(base + o_hi), (const int32_t *)(base + o_ptr),
        (const float *)(base + o_w), (int)w.size(), p->log_db, logmel, umax);
    SAPR_LAUNCH_CHECK(ctx);
    const int64_t n = total_frames * n_mfcc;
    k_mfcc_dct<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(logmel, uof, umax, n_mels,
                                                                    n_mfcc, p->log_db ? p->top_db : 0.0f,
                                                                    (const float *)(base + o_dct), feats, ld_out,
                                                                    total_frames);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}
