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
// k_mfcc_logmel: one CTA per frame: samples (+pre-emphasis, window) -> shared memory, in-place
//                radix-2 FFT in shared memory, power spectrum, sparse triangular mel filters,
//                log -> logmel[frame][n_mels] and the per-utterance maximum (ordered-int atomicMax).
// k_mfcc_dct:    clamp to max - top_db, DCT-II, write feats[frame][ld_out].
#include "common.cuh"

#define MFCC_PI 3.14159265358979323846

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

__global__ void k_mfcc_logmel(const float *__restrict__ audio, const int64_t *__restrict__ sample_off,
                              const int64_t *__restrict__ frame_off, const int32_t *__restrict__ utt_of_frame,
                              int n_fft, int log2n, int hop, int center, float preemph,
                              const float *__restrict__ window, const float2 *__restrict__ twiddle, int n_mels,
                              const int32_t *__restrict__ mel_lo, const int32_t *__restrict__ mel_hi,
                              const int32_t *__restrict__ mel_ptr, const float *__restrict__ mel_w, int log_db,
                              float *__restrict__ logmel, int *__restrict__ umax) {
    extern __shared__ float2 s_x[];   // n_fft complex, then n_fft/2+1 power values overlay
    const int64_t f = blockIdx.x;
    const int u = utt_of_frame[f];
    const int64_t s0 = sample_off[u], len = sample_off[u + 1] - s0;
    const int64_t fi = f - frame_off[u];
    const int64_t start = fi * hop - (center ? n_fft / 2 : 0);
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
        const int64_t s = start + i;
        float v = 0.0f;
        if (s >= 0 && s < len) {
            v = audio[s0 + s];
            if (preemph != 0.0f && s > 0) v -= preemph * audio[s0 + s - 1];
        }
        v *= window[i];
        const int r = __brev((unsigned)i) >> (32 - log2n);
        s_x[r] = make_float2(v, 0.0f);
    }
    __syncthreads();
    for (int st = 1; st <= log2n; st++) {
        const int half = 1 << (st - 1);
        for (int k = threadIdx.x; k < n_fft / 2; k += blockDim.x) {
            const int grp = k / half, pos = k % half;
            const int i0 = grp * 2 * half + pos, i1 = i0 + half;
            const float2 w = twiddle[pos * (n_fft / (2 * half))];
            const float2 a = s_x[i0], b = s_x[i1];
            const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            s_x[i0] = make_float2(a.x + t.x, a.y + t.y);
            s_x[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    float *pw = reinterpret_cast<float *>(s_x);   // power spectrum overlays the first half of the buffer
    const int nb = n_fft / 2 + 1;
    // pw[k] aliases s_x[k/2]: thread k writes float index k after reading complex index k >= k/2, so go
    // through registers and a barrier
    float pv[16];
    int cnt = 0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) { const float2 c = s_x[k]; pv[cnt++] = c.x * c.x + c.y * c.y; }
    __syncthreads();
    cnt = 0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) pw[k] = pv[cnt++];
    __syncthreads();
    for (int mth = threadIdx.x; mth < n_mels; mth += blockDim.x) {
        float acc = 0.0f;
        const float *w = mel_w + mel_ptr[mth];
        for (int k = mel_lo[mth]; k < mel_hi[mth]; k++) acc += w[k - mel_lo[mth]] * pw[k];
        acc = fmaxf(acc, 1e-10f);
        const float lv = log_db ? 10.0f * log10f(acc) : logf(acc);
        logmel[f * n_mels + mth] = lv;
        atomicMax(umax + u, float_to_ordered(lv));
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
    std::vector<int32_t> uof(total_frames);
    for (int u = 0; u < B; u++) for (int64_t f = foff[u]; f < foff[u + 1]; f++) uof[f] = u;
    // ---- device tables: one workspace blob ----
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
#define UP(off, vec, T) SAPR_CUDA(ctx, cudaMemcpyAsync(base + off, vec.data(), sizeof(T) * vec.size(), cudaMemcpyHostToDevice, ctx->stream))
    UP(o_win, window, float); UP(o_tw, tw, float2); UP(o_lo, lo, int32_t); UP(o_hi, hi, int32_t); UP(o_ptr, ptr, int32_t);
    UP(o_w, w, float); UP(o_dct, dct, float); UP(o_foff, foff, int64_t); UP(o_uof, uof, int32_t);
#undef UP
    SAPR_CUDA(ctx, cudaMemcpyAsync(base + o_soff, sample_offsets_host, sizeof(int64_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream));
    SAPR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the std::vectors above go out of scope
    int *umax = (int *)(base + o_umax);
    float *logmel = (float *)(base + o_lm);
    k_mfcc_init_max<<<(B + 127) / 128, 128, 0, ctx->stream>>>(umax, B);
    SAPR_LAUNCH_CHECK(ctx);
    const int threads = std::min(256, n_fft / 2);
    if ((nb + threads - 1) / threads > 16) SAPR_FAIL(ctx, SAPR_E_RANGE, "mfcc: n_fft too large for the register staging");
    k_mfcc_logmel<<<(unsigned)total_frames, threads, sizeof(float2) * n_fft, ctx->stream>>>(
        audio, (const int64_t *)(base + o_soff), (const int64_t *)(base + o_foff), (const int32_t *)(base + o_uof), n_fft,
        log2n, p->hop_length, p->center, p->preemph, (const float *)(base + o_win), (const float2 *)(base + o_tw), n_mels,
        (const int32_t *)(base + o_lo), (const int32_t *)(base + o_hi), (const int32_t *)(base + o_ptr),
        (const float *)(base + o_w), p->log_db, logmel, umax);
    SAPR_LAUNCH_CHECK(ctx);
    const int64_t n = total_frames * n_mfcc;
    k_mfcc_dct<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(logmel, (const int32_t *)(base + o_uof), umax, n_mels,
                                                                    n_mfcc, p->log_db ? p->top_db : 0.0f,
                                                                    (const float *)(base + o_dct), feats, ld_out,
                                                                    total_frames);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}
