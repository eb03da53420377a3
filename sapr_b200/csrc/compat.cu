// compat.cu -- float64 kernels on MATERIALISED matrices: the per-step methods of class HMM
// (custom_hmm.py), kept so that the reference's own tests and the as-written (sapr) semantics run on
// the CUDA path: compute_emission_matrix (:146-174, Gram row-sum / full covariance),
// forward (:176-211), backward (:213-246), compute_gamma (:248-257), compute_xi (:259-322),
// update_A (:351-364), update_B (:366-400), decode (:462-514).  These run one thread per utterance
// (the recursions are sequential in t) -- cfg-1 sized work is launch/latency bound by nature; the
// throughput path is the fused kernels in viterbi.cu / estep.cu.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// emission
// DIAG: E[f][j] for one model, all frames of a batch.  grid over frames, blockDim.x = 128.
__global__ void k_emission_diag_mat(const float *__restrict__ X, int ldx, int64_t total_frames, int D, int S,
                                    int all_emit, const double *__restrict__ mean, const double *__restrict__ var,
                                    double *__restrict__ E) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_frames * S) return;
    const int64_t f = idx / S;
    const int j = (int)(idx % S);
    if (!all_emit && (j == 0 || j == S - 1)) { E[idx] = -INFINITY; return; }
    const double *mu = mean + (size_t)j * D, *vr = var + (size_t)j * D;
    double ld = 0.0, q = 0.0;
    for (int d = 0; d < D; d++) {
        ld += log(vr[d]);
        const double df = (double)X[f * ldx + d] - mu[d];
        q += df * df / vr[d];
    }
    E[idx] = -0.5 * (D * SAPR_LOG2PI + ld + q);
}

// SAPR (D1): per (utterance, state): s = sum_tau (x_tau - mu_j), v = P_j s.  grid = (B, N), 64 threads.
__global__ void k_sapr_v(const float *__restrict__ X, int ldx, const int64_t *__restrict__ offsets, int D, int S,
                         const double *__restrict__ mean, const double *__restrict__ P, double *__restrict__ v) {
    const int u = blockIdx.x, j = blockIdx.y + 1;
    __shared__ double s_s[64];
    const int64_t off = offsets[u];
    const int T = (int)(offsets[u + 1] - off);
    const int d = threadIdx.x;
    if (d < D) {
        const double mu = mean[(size_t)j * D + d];
        double s = 0.0;
        for (int t = 0; t < T; t++) s += (double)X[(off + t) * ldx + d] - mu;
        s_s[d] = s;
    }
    __syncthreads();
    if (d < D) {
        const double *Pj = P + (size_t)j * D * D;
        double acc = 0.0;
        for (int b = 0; b < D; b++) acc += Pj[(size_t)d * D + b] * s_s[b];
        v[((size_t)u * S + j) * D + d] = acc;
    }
}

__global__ void k_emission_sapr_mat(const float *__restrict__ X, int ldx, const int64_t *__restrict__ offsets, int B,
                                    int D, int S, const double *__restrict__ mean, const double *__restrict__ cstS,
                                    const double *__restrict__ v, const int32_t *__restrict__ utt_of_frame,
                                    int64_t total_frames, double *__restrict__ E) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_frames * S) return;
    const int64_t f = idx / S;
    const int j = (int)(idx % S);
    if (j == 0 || j == S - 1) { E[idx] = -INFINITY; return; }
    const int u = utt_of_frame ? utt_of_frame[f] : 0;
    const double *mu = mean + (size_t)j * D;
    const double *vj = v + ((size_t)u * S + j) * D;
    double q = 0.0;
    for (int d = 0; d < D; d++) q += ((double)X[f * ldx + d] - mu[d]) * vj[d];
    // -0.5 (D ln 2pi + logdet + q), with cstS = -0.5 (D ln 2pi + logdet)
    E[idx] = cstS[j] - 0.5 * q;
}

__global__ void k_utt_of_frame(const int64_t *__restrict__ offsets, int B, int32_t *__restrict__ uof) {
    const int u = blockIdx.x;
    for (int64_t f = offsets[u] + threadIdx.x; f < offsets[u + 1]; f += blockDim.x) uof[f] = u;
}

// ---------------------------------------------------------------------------------------------
// device restatements of the recursions (same control flow as the reference loops)
__device__ void d_forward(const double *E, int T, int S, const double *A, double *alpha, double *scale_out) {
    const double NINF = -INFINITY;
    for (int i = 0; i < T * S; i++) alpha[i] = NINF;
    alpha[0] = 0.0;
    alpha[1] = log(A[0 * S + 1]) + E[1];
    for (int t = 1; t < T; t++) {
        const double *ap = alpha + (size_t)(t - 1) * S;
        double *ac = alpha + (size_t)t * S;
        ac[0] = NINF;
        for (int j = 1; j < S; j++) {
            if (j == 1) ac[j] = lae(ap[0] + log(A[0 * S + 1]), ap[1] + log(A[1 * S + 1])) + E[(size_t)t * S + j];
            else if (j < S - 1) ac[j] = lae(ap[j - 1] + log(A[(j - 1) * S + j]), ap[j] + log(A[j * S + j])) + E[(size_t)t * S + j];
            else ac[j] = ap[j - 1] + log(A[(j - 1) * S + j]);
        }
    }
    double mx = alpha[0];
    bool has_nan = false;
    for (int i = 0; i < T * S; i++) {
        if (alpha[i] != alpha[i]) has_nan = true;
        else if (alpha[i] > mx) mx = alpha[i];
    }
    if (has_nan) mx = NAN;
    for (int i = 0; i < T * S; i++) alpha[i] -= mx;
    *scale_out = mx;
}

__device__ void d_backward(const double *E, int T, int S, const double *A, double scale, double *beta) {
    const double NINF = -INFINITY;
    for (int i = 0; i < T * S; i++) beta[i] = NINF;
    beta[(size_t)(T - 1) * S + S - 1] = 0.0;
    for (int t = T - 2; t >= 0; t--) {
        const double *bn = beta + (size_t)(t + 1) * S;
        const double *en = E + (size_t)(t + 1) * S;
        double *bc = beta + (size_t)t * S;
        for (int i = 0; i < S - 1; i++) {
            if (i == 0) bc[i] = log(A[0 * S + 1]) + en[1] + bn[1];
            else if (i < S - 2)
                bc[i] = lae(log(A[i * S + i]) + en[i] + bn[i], log(A[i * S + i + 1]) + en[i + 1] + bn[i + 1]);
            else
                bc[i] = lae(log(A[i * S + i]) + en[i] + bn[i], log(A[i * S + i + 1]) + bn[i + 1]);
        }
    }
    for (int i = 0; i < (T - 1) * S; i++) beta[i] -= scale;
}

__device__ double d_lae_reduce(const double *v, int n) {
    double r = v[0];
    for (int i = 1; i < n; i++) r = lae(r, v[i]);
    return r;
}

__device__ void d_gamma(const double *alpha, const double *beta, int T, int S, double *gamma) {
    for (int t = 0; t < T; t++) {
        double r = alpha[(size_t)t * S] + beta[(size_t)t * S];
        for (int j = 1; j < S; j++) r = lae(r, alpha[(size_t)t * S + j] + beta[(size_t)t * S + j]);
        for (int j = 0; j < S; j++) gamma[(size_t)t * S + j] = exp(alpha[(size_t)t * S + j] + beta[(size_t)t * S + j] - r);
    }
}

// xi row t into x[S*S] (custom_hmm.py:270-320)
__device__ void d_xi_row(const double *alpha, const double *beta, const double *E, int t, int S, const double *A,
                         double ll, double *x) {
    for (int k = 0; k < S * S; k++) x[k] = 0.0;
    const double *a = alpha + (size_t)t * S;
    const double *bn = beta + (size_t)(t + 1) * S;
    const double *en = E + (size_t)(t + 1) * S;
    x[0 * S + 1] = exp(a[0] + log(A[0 * S + 1]) + en[1] + bn[1] - ll);
    for (int i = 1; i < S - 1; i++) {
        if (A[i * S + i] > 0) x[i * S + i] = exp(a[i] + log(A[i * S + i]) + en[i] + bn[i] - ll);
        if (i < S - 2) x[i * S + i + 1] = exp(a[i] + log(A[i * S + i + 1]) + en[i + 1] + bn[i + 1] - ll);
    }
    x[(S - 2) * S + (S - 1)] = exp(a[S - 2] + log(A[(S - 2) * S + S - 1]) + en[S - 1] + bn[S - 1] - ll);
    x[(S - 1) * S + (S - 1)] = exp(a[S - 1] + log(A[(S - 1) * S + S - 1]) + en[S - 1] + bn[S - 1] - ll);
    double sum = 0.0;
    for (int k = 0; k < S * S; k++) sum += x[k];
    if (sum > 0)
        for (int k = 0; k < S * S; k++) x[k] /= sum;
}

__device__ double d_decode(const double *E, int T_eff, int S, const double *A, double *V, int32_t *bp, int32_t *path) {
    const int N = S - 2;
    const double NINF = -INFINITY;
    for (int i = 0; i < T_eff * S; i++) { V[i] = NINF; bp[i] = 0; }
    V[0] = 0.0;
    V[1] = log(A[0 * S + 1]) + E[1];
    for (int t = 1; t < T_eff; t++)
        for (int j = 1; j < S; j++) {
            int cand[2], nc = 0;
            if (j == 1) { cand[nc++] = 1; if (t == 1) cand[nc++] = 0; }
            else if (j == S - 1) { if (t >= N) { cand[nc++] = j - 1; cand[nc++] = j; } else continue; }
            else { cand[nc++] = j - 1; cand[nc++] = j; }
            double best = NINF;
            int best_prev = -1;
            for (int c = 0; c < nc; c++) {
                const int i = cand[c];
                const double score = V[(size_t)(t - 1) * S + i] + log(A[i * S + j]);
                if (score > best) { best = score; best_prev = i; }
            }
            if (best_prev >= 0) {
                V[(size_t)t * S + j] = (j != S - 1) ? best + E[(size_t)t * S + j] : best;
                bp[(size_t)t * S + j] = best_prev;
            }
        }
    int cur = S - 1;
    for (int t = T_eff - 1; t >= 0; t--) { path[t] = cur; cur = bp[(size_t)t * S + cur]; }
    return V[(size_t)(T_eff - 1) * S + S - 1];
}

// ---------------------------------------------------------------------------------------------
// single-utterance API kernels
__global__ void k_forward1(const double *E, int T, int S, const double *A, double *alpha, double *scale) {
    if (blockIdx.x == 0 && threadIdx.x == 0) d_forward(E, T, S, A, alpha, scale);
}
__global__ void k_backward1(const double *E, int T, int S, const double *A, const double *scale, double *beta) {
    if (blockIdx.x == 0 && threadIdx.x == 0) d_backward(E, T, S, A, *scale, beta);
}
__global__ void k_gamma1(const double *alpha, const double *beta, int T, int S, double *gamma) {
    if (blockIdx.x == 0 && threadIdx.x == 0) d_gamma(alpha, beta, T, S, gamma);
}
__global__ void k_xi1(const double *alpha, const double *beta, const double *E, int T, int S, const double *A, double *xi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T - 1) return;
    const double ll = d_lae_reduce(alpha + (size_t)(T - 1) * S, S);
    d_xi_row(alpha, beta, E, t, S, A, ll, xi + (size_t)t * S * S);
}
__global__ void k_decode1(const double *E, int T_eff, int S, const double *A, double *V, int32_t *bp, double *score,
                          int32_t *path) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *score = d_decode(E, T_eff, S, A, V, bp, path);
}

// batched E-step, one thread per utterance: alpha/beta/xi-row scratch in global memory
__global__ void k_estep_compat(const double *__restrict__ E, const int64_t *__restrict__ offsets, int B, int S,
                               const double *__restrict__ A, double *__restrict__ alpha, double *__restrict__ beta,
                               double *__restrict__ gamma, double *__restrict__ xirow, double *__restrict__ per_utt,
                               double *__restrict__ loglik) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    const int64_t off = offsets[u];
    const int T = (int)(offsets[u + 1] - off);
    const double *Eu = E + (size_t)off * S;
    double *al = alpha + (size_t)off * S, *be = beta + (size_t)off * S, *ga = gamma + (size_t)off * S;
    double *xr = xirow + (size_t)u * S * S;
    double *pu = per_utt + (size_t)u * (S + S * S);
    for (int k = 0; k < S + S * S; k++) pu[k] = 0.0;
    if (T <= 0) { loglik[u] = 0.0; return; }
    double scale;
    d_forward(Eu, T, S, A, al, &scale);
    d_backward(Eu, T, S, A, scale, be);
    d_gamma(al, be, T, S, ga);
    for (int t = 0; t < T - 1; t++)
        for (int j = 0; j < S; j++) pu[j] += ga[(size_t)t * S + j];
    const double ll = d_lae_reduce(al + (size_t)(T - 1) * S, S);
    for (int t = 0; t < T - 1; t++) {
        d_xi_row(al, be, Eu, t, S, A, ll, xr);
        for (int k = 0; k < S * S; k++) pu[S + k] += xr[k];
    }
    loglik[u] = ll;
}

__global__ void k_sum_per_utt(const double *__restrict__ per_utt, int B, int len, double *__restrict__ agg_gamma, int S,
                              double *__restrict__ agg_xi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    double s = 0.0;
    for (int u = 0; u < B; u++) s += per_utt[(size_t)u * len + k];   // utterance order, like the reference loop
    if (k < S) agg_gamma[k] = s; else agg_xi[k - S] = s;
}

// update_A (custom_hmm.py:351-364)
__global__ void k_update_A(int S, const double *__restrict__ agg_xi, const double *__restrict__ agg_gamma, double *A) {
    if (threadIdx.x == 0) { A[0 * S + 1] = 1.0; A[(size_t)(S - 1) * S + S - 1] = 1.0; }
    for (int i = threadIdx.x; i < S - 1; i += blockDim.x) {   // any number of states
        if (i < 1 || !(agg_gamma[i] > 0)) continue;
        const double aii = agg_xi[(size_t)i * S + i] / agg_gamma[i];
        A[(size_t)i * S + i] = aii;
        A[(size_t)i * S + i + 1] = 1.0 - aii;
    }
}

// update_B pass 1 (custom_hmm.py:372-379): one thread per (state, dim)
__global__ void k_update_B_mean(const float *__restrict__ X, int ldx, int64_t total_frames, int D, int S,
                                const double *__restrict__ gamma, double *__restrict__ mean, double *__restrict__ occ) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * D) return;
    const int j = idx / D, d = idx % D;
    double s = 0.0, o = 0.0;
    if (j >= 1 && j < S - 1) {
        for (int64_t f = 0; f < total_frames; f++) {
            const double g = gamma[f * S + j];
            s += g * (double)X[f * ldx + d];
            o += g;
        }
        if (o > 0) s /= o;
    }
    mean[idx] = s;
    if (d == 0) occ[j] = o;
}

// update_B pass 2 (:382-397): one thread per (state, a, b); SAPR models keep the FULL covariance (D2),
// DIAG models only the diagonal.
__global__ void k_update_B_cov(const float *__restrict__ X, int ldx, int64_t total_frames, int D, int S, int full,
                               const double *__restrict__ gamma, const double *__restrict__ mean,
                               const double *__restrict__ occ, double floor_var, double *__restrict__ cov) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = full ? D * D : D;
    if (idx >= S * per) return;
    const int j = idx / per, r = idx % per;
    const int a = full ? r / D : r, b = full ? r % D : r;
    double c = 0.0;
    if (j >= 1 && j < S - 1) {
        const double ma = mean[(size_t)j * D + a], mb = mean[(size_t)j * D + b];
        for (int64_t f = 0; f < total_frames; f++) {
            const double g = gamma[f * S + j];
            const double da = (double)X[f * ldx + a] - ma, db = (double)X[f * ldx + b] - mb;
            c += g * (da * db);
        }
        const double o = occ[j];
        if (o > 0) {
            c /= o;
            if (a == b) c = (c != c) ? c : (c > floor_var ? c : floor_var);
        }
    }
    cov[idx] = c;
}

// ---------------------------------------------------------------------------------------------
// host entry points
int sapr_emission_into(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, const int64_t *offsets, int B,
                         int64_t total_frames, double *E) {
    const int S = m->S, D = m->D;
    const int64_t n = total_frames * S;
    if (n <= 0) return SAPR_OK;
    const double *mean = m->mean + (size_t)mi * S * D;
    if (m->emission == SAPR_EMIT_DIAG) {
        const double *var = m->cov + (size_t)mi * S * D;
        k_emission_diag_mat<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
            X, ldx, total_frames, D, S, m->topology == SAPR_TOPO_DENSE, mean, var, E);
        SAPR_LAUNCH_CHECK(ctx);
        return SAPR_OK;
    }
    // SAPR: v[B][S][D] + utterance-of-frame map
    int rc = sapr_ws_reserve(ctx, 7, sizeof(double) * (size_t)B * S * D + sizeof(int32_t) * (size_t)total_frames + 64);
    if (rc) return rc;
    double *v = (double *)ctx->ws[7];
    int32_t *uof = (int32_t *)(v + (size_t)B * S * D);
    const double *P = m->P + (size_t)mi * S * D * D;
    const double *cstS = m->cstS + (size_t)mi * S;
    k_sapr_v<<<dim3(B, m->N), 64, 0, ctx->stream>>>(X, ldx, offsets, D, S, mean, P, v);
    SAPR_LAUNCH_CHECK(ctx);
    k_utt_of_frame<<<B, 64, 0, ctx->stream>>>(offsets, B, uof);
    SAPR_LAUNCH_CHECK(ctx);
    k_emission_sapr_mat<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(X, ldx, offsets, B, D, S, mean, cstS, v, uof,
                                                                              total_frames, E);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

// offsets for a single utterance of T frames, kept in workspace slot 5
static int single_offsets(sapr_ctx *ctx, int T, int64_t **out) {
    int rc = sapr_ws_reserve(ctx, 5, 64);
    if (rc) return rc;
    int64_t h[2] = {0, T};
    SAPR_CUDA(ctx, cudaMemcpyAsync(ctx->ws[5], h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    SAPR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = (int64_t *)ctx->ws[5];
    return SAPR_OK;
}

#define CHECK_MODEL(fn)                                                                  \
    if (!ctx || !m) return SAPR_E_INVALID;                                               \
    if (!m->valid) SAPR_FAIL(ctx, SAPR_E_INVALID, fn ": model parameters not set");      \
    if (mi < 0 || mi >= m->M) SAPR_FAIL(ctx, SAPR_E_INVALID, fn ": model index out of range")

extern "C" int sapr_emission(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, int T, double *E) {
    CHECK_MODEL("emission");
    if (!X || !E || T <= 0) return SAPR_E_INVALID;
    int64_t *offs;
    int rc = single_offsets(ctx, T, &offs);
    if (rc) return rc;
    return sapr_emission_into(ctx, m, mi, X, ldx, offs, 1, T, E);
}

extern "C" int sapr_forward(sapr_ctx *ctx, sapr_models *m, int mi, const double *E, int T, double *alpha, double *scale) {
    CHECK_MODEL("forward");
    if (m->topology != SAPR_TOPO_ENTRY_EXIT) SAPR_FAIL(ctx, SAPR_E_INVALID, "forward: ENTRY_EXIT topology only");
    k_forward1<<<1, 32, 0, ctx->stream>>>(E, T, m->S, m->A + (size_t)mi * m->S * m->S, alpha, scale);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

extern "C" int sapr_backward(sapr_ctx *ctx, sapr_models *m, int mi, const double *E, int T, const double *scale, double *beta) {
    CHECK_MODEL("backward");
    if (m->topology != SAPR_TOPO_ENTRY_EXIT) SAPR_FAIL(ctx, SAPR_E_INVALID, "backward: ENTRY_EXIT topology only");
    k_backward1<<<1, 32, 0, ctx->stream>>>(E, T, m->S, m->A + (size_t)mi * m->S * m->S, scale, beta);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

extern "C" int sapr_gamma(sapr_ctx *ctx, int S, const double *alpha, const double *beta, int T, double *gamma) {
    if (!ctx || !alpha || !beta || !gamma) return SAPR_E_INVALID;
    k_gamma1<<<1, 32, 0, ctx->stream>>>(alpha, beta, T, S, gamma);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

extern "C" int sapr_xi(sapr_ctx *ctx, sapr_models *m, int mi, const double *alpha, const double *beta, const double *E,
                       int T, double *xi) {
    CHECK_MODEL("xi");
    if (T <= 1) return SAPR_OK;
    k_xi1<<<(T - 1 + 63) / 64, 64, 0, ctx->stream>>>(alpha, beta, E, T, m->S, m->A + (size_t)mi * m->S * m->S, xi);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

extern "C" int sapr_decode_mat(sapr_ctx *ctx, sapr_models *m, int mi, const double *E, int T_eff, double *score, int32_t *path) {
    CHECK_MODEL("decode_mat");
    if (T_eff <= 0) return SAPR_E_INVALID;
    const int S = m->S;
    int rc = sapr_ws_reserve(ctx, 6, (size_t)T_eff * S * (sizeof(double) + sizeof(int32_t)));
    if (rc) return rc;
    double *V = (double *)ctx->ws[6];
    int32_t *bp = (int32_t *)(V + (size_t)T_eff * S);
    k_decode1<<<1, 32, 0, ctx->stream>>>(E, T_eff, S, m->A + (size_t)mi * S * S, V, bp, score, path);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}

extern "C" int sapr_decode_compat(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, int T, int T_eff,
                                  double *score, int32_t *path) {
    CHECK_MODEL("decode_compat");
    if (T <= 0 || T_eff <= 0) return SAPR_E_INVALID;
    if (T_eff > T) SAPR_FAIL(ctx, SAPR_E_SHORT, "decode: utterance shorter than the frames to walk");
    int rc = sapr_ws_reserve(ctx, 4, sizeof(double) * (size_t)T * m->S);
    if (rc) return rc;
    double *E = (double *)ctx->ws[4];
    int64_t *offs;
    if ((rc = single_offsets(ctx, T, &offs))) return rc;
    if ((rc = sapr_emission_into(ctx, m, mi, X, ldx, offs, 1, T, E))) return rc;
    return sapr_decode_mat(ctx, m, mi, E, T_eff, score, path);
}

extern "C" int sapr_update_A(sapr_ctx *ctx, sapr_models *m, int mi, const double *agg_xi, const double *agg_gamma) {
    CHECK_MODEL("update_A");
    k_update_A<<<1, 64, 0, ctx->stream>>>(m->S, agg_xi, agg_gamma, m->A + (size_t)mi * m->S * m->S);
    SAPR_LAUNCH_CHECK(ctx);
    return sapr_models_prepare(m);
}

extern "C" int sapr_update_B(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, int64_t total_frames,
                             const double *gamma, double floor_var) {
    CHECK_MODEL("update_B");
    const int S = m->S, D = m->D;
    int rc = sapr_ws_reserve(ctx, 5, sizeof(double) * (S + 8));
    if (rc) return rc;
    double *occ = (double *)ctx->ws[5];
    double *mean = m->mean + (size_t)mi * S * D;
    const int full = (m->emission == SAPR_EMIT_SAPR);
    double *cov = m->cov + (size_t)mi * S * (full ? D * D : D);
    k_update_B_mean<<<(S * D + 63) / 64, 64, 0, ctx->stream>>>(X, ldx, total_frames, D, S, gamma, mean, occ);
    SAPR_LAUNCH_CHECK(ctx);
    const int n = S * (full ? D * D : D);
    k_update_B_cov<<<(n + 63) / 64, 64, 0, ctx->stream>>>(X, ldx, total_frames, D, S, full, gamma, mean, occ, floor_var, cov);
    SAPR_LAUNCH_CHECK(ctx);
    return sapr_models_prepare(m);
}

extern "C" int sapr_estep_compat(sapr_ctx *ctx, sapr_models *m, int mi, const float *X, int ldx, const int64_t *offsets,
                                 int B, int64_t total_frames, double *gamma, double *agg_gamma, double *agg_xi,
                                 double *loglik) {
    CHECK_MODEL("estep_compat");
    if (m->topology != SAPR_TOPO_ENTRY_EXIT) SAPR_FAIL(ctx, SAPR_E_INVALID, "estep_compat: ENTRY_EXIT topology only");
    const int S = m->S;
    const size_t lat = (size_t)total_frames * S;
    const size_t need = sizeof(double) * (3 * lat + (size_t)B * S * S + (size_t)B * (S + S * S));
    int rc = sapr_ws_reserve(ctx, 4, need);
    if (rc) return rc;
    double *E = (double *)ctx->ws[4];
    double *alpha = E + lat, *beta = alpha + lat, *xirow = beta + lat, *per_utt = xirow + (size_t)B * S * S;
    if ((rc = sapr_emission_into(ctx, m, mi, X, ldx, offsets, B, total_frames, E))) return rc;
    k_estep_compat<<<(B + 31) / 32, 32, 0, ctx->stream>>>(E, offsets, B, S, m->A + (size_t)mi * S * S, alpha, beta, gamma,
                                                          xirow, per_utt, loglik);
    SAPR_LAUNCH_CHECK(ctx);
    const int len = S + S * S;
    k_sum_per_utt<<<(len + 63) / 64, 64, 0, ctx->stream>>>(per_utt, B, len, agg_gamma, S, agg_xi);
    SAPR_LAUNCH_CHECK(ctx);
    return SAPR_OK;
}
