/*
 * sapr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A float64 CPU restatement of the reference's assignment2 HMM hot path
 * (frankcholula/sapr, assignment2/custom_hmm.py), written in plain C so that it
 * travels to the GPU box and finishes in seconds.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load this
 * library; nothing under sapr_b200/ imports it.
 *
 * Parity status: PINNED for the custom_hmm.py path -- tests/test_oracle_golden.py
 * checks every function below against vectors produced by importing the
 * reference's own custom_hmm.HMM in the build container
 * (tools/make_golden.py -> tests/golden/ npz files).  The hmmlearn-style functions
 * (orc_hl_*) restate hmmlearn 0.3.3 (assignment2/poetry.lock:430-431), which is
 * NOT vendored in the reference and not installable here: parity unpinned.
 *
 * Layout conventions: features are frame-major X[T][D] doubles (the reference
 * holds (D,T) arrays; the Python wrapper transposes).  S = N + 2 total states,
 * state 0 = entry, state S-1 = exit.  All matrices row-major.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_LOG2PI 1.8378770664093454835606594728112 /* ln(2*pi) */
#define ORC_LN2 0.69314718055994530941723212145818

/* numpy's npy_logaddexp (numpy/_core/src/npymath/npy_math_internal.h.src), the
 * function behind every np.logaddexp call in custom_hmm.py:190,198,228,237. */
static double orc_logaddexp(double x, double y) {
    if (x == y) return x + ORC_LN2; /* handles (-inf,-inf) -> -inf and (inf,inf) */
    double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    else if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp; /* NaN */
}

/* np.logaddexp.reduce over a vector: sequential left fold. */
static double orc_logaddexp_reduce(const double *v, int n) {
    double r = v[0];
    for (int i = 1; i < n; i++) r = orc_logaddexp(r, v[i]);
    return r;
}

/* LU factorisation with partial pivoting (what LAPACK getrf does for
 * np.linalg.inv / np.linalg.slogdet, custom_hmm.py:164-165).  a is n*n and is
 * overwritten; returns log|det| and writes the inverse into inv. */
static double orc_lu_inverse(double *a, int n, double *inv) {
    int *piv = (int *)malloc(sizeof(int) * n);
    double logabsdet = 0.0;
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = fabs(a[k * n + k]);
        for (int i = k + 1; i < n; i++) {
            double v = fabs(a[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; j++) {
                double t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t;
            }
        double d = a[k * n + k];
        logabsdet += log(fabs(d));
        for (int i = k + 1; i < n; i++) {
            a[i * n + k] /= d;
            double l = a[i * n + k];
            for (int j = k + 1; j < n; j++) a[i * n + j] -= l * a[k * n + j];
        }
    }
    /* solve A X = I column by column: apply P, forward (unit L), backward (U) */
    double *col = (double *)malloc(sizeof(double) * n);
    for (int c = 0; c < n; c++) {
        for (int i = 0; i < n; i++) col[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; k++)
            if (piv[k] != k) { double t = col[k]; col[k] = col[piv[k]]; col[piv[k]] = t; }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++) col[i] -= a[i * n + j] * col[j];
        for (int i = n - 1; i >= 0; i--) {
            for (int j = i + 1; j < n; j++) col[i] -= a[i * n + j] * col[j];
            col[i] /= a[i * n + i];
        }
        for (int i = 0; i < n; i++) inv[i * n + c] = col[i];
    }
    free(col);
    free(piv);
    return logabsdet;
}

/* ------------------------------------------------------------------------- */
/* custom_hmm.py:35-116 -- flat start.  X is the concatenation of all B
 * utterances, offsets[B+1] in frames.                                        */
void orc_init_parameters(const double *X, const int64_t *offsets, int B, int D, int N,
                         double var_floor_factor, double *global_mean, double *global_cov,
                         double *A, double *mean, double *cov) {
    const int S = N + 2;
    int64_t total = offsets[B];
    /* calculate_means (:70-80) */
    for (int d = 0; d < D; d++) global_mean[d] = 0.0;
    for (int64_t t = 0; t < total; t++)
        for (int d = 0; d < D; d++) global_mean[d] += X[t * D + d];
    for (int d = 0; d < D; d++) global_mean[d] /= (double)total;
    /* calculate_covariance (:82-92), then off-diagonals zeroed (:42) */
    memset(global_cov, 0, sizeof(double) * D * D);
    for (int64_t t = 0; t < total; t++)
        for (int d = 0; d < D; d++) {
            double c = X[t * D + d] - global_mean[d];
            global_cov[d * D + d] += c * c;
        }
    double tr = 0.0;
    for (int d = 0; d < D; d++) { global_cov[d * D + d] /= (double)total; tr += global_cov[d * D + d]; }
    /* variance floor (:45-49) */
    double floor_v = var_floor_factor * (tr / D);
    for (int d = 0; d < D; d++)
        if (global_cov[d * D + d] < floor_v) global_cov[d * D + d] = floor_v;
    /* initialize_transitions (:94-116) */
    double avg = (double)total / ((double)B * (double)N);
    double aii = exp(-1.0 / (avg - 1.0));
    memset(A, 0, sizeof(double) * S * S);
    A[0 * S + 1] = 1.0;
    for (int i = 1; i <= N; i++) { A[i * S + i] = aii; A[i * S + i + 1] = 1.0 - aii; }
    A[(S - 1) * S + (S - 1)] = 1.0;
    /* B (:51-61): every state (entry/exit included) gets the global statistics */
    for (int s = 0; s < S; s++) {
        memcpy(mean + (size_t)s * D, global_mean, sizeof(double) * D);
        memcpy(cov + (size_t)s * D * D, global_cov, sizeof(double) * D * D);
    }
}

/* custom_hmm.py:146-174 AS WRITTEN: the quadratic term is the row
 * sum of the T x T Gram matrix diff^T P diff, not its diagonal.               */
void orc_emission_sapr(const double *X, int T, int D, int S, const double *mean,
                       const double *cov, double *E) {
    double *c = (double *)malloc(sizeof(double) * D * D);
    double *P = (double *)malloc(sizeof(double) * D * D);
    double *diff = (double *)malloc(sizeof(double) * (size_t)T * D);
    double *u = (double *)malloc(sizeof(double) * D);
    for (int t = 0; t < T; t++) { E[(size_t)t * S] = -INFINITY; E[(size_t)t * S + S - 1] = -INFINITY; }
    for (int j = 1; j < S - 1; j++) {
        for (int a = 0; a < D; a++)
            for (int b = 0; b < D; b++)
                c[a * D + b] = cov[((size_t)j * D + a) * D + b] + (a == b ? 1e-6 : 0.0);
        double logdet = orc_lu_inverse(c, D, P);
        for (int t = 0; t < T; t++)
            for (int d = 0; d < D; d++) diff[(size_t)t * D + d] = X[(size_t)t * D + d] - mean[(size_t)j * D + d];
        for (int t = 0; t < T; t++) {
            /* u = d_t^T P  (row t of diff.T @ inv_cov) */
            for (int b = 0; b < D; b++) {
                double s = 0.0;
                for (int a = 0; a < D; a++) s += diff[(size_t)t * D + a] * P[a * D + b];
                u[b] = s;
            }
            /* sum_tau u . d_tau  (np.sum(..., axis=1) of the Gram row) */
            double q = 0.0;
            for (int tau = 0; tau < T; tau++) {
                double g = 0.0;
                for (int b = 0; b < D; b++) g += u[b] * diff[(size_t)tau * D + b];
                q += g;
            }
            E[(size_t)t * S + j] = -0.5 * (D * ORC_LOG2PI + logdet + q);
        }
    }
    free(u); free(diff); free(P); free(c);
}

/* "standard" emission: true diagonal Gaussian log-density (SURVEY App. A;
 * the Rung-1 oracle overrides only this).  var is S x D.  If all_emit == 0 the
 * entry/exit columns are -inf (sapr topology), else every state emits
 * (hmmlearn topology).                                                       */
void orc_emission_diag(const double *X, int T, int D, int S, const double *mean,
                       const double *var, int all_emit, double *E) {
    for (int j = 0; j < S; j++) {
        int emits = all_emit || (j > 0 && j < S - 1);
        if (!emits) { for (int t = 0; t < T; t++) E[(size_t)t * S + j] = -INFINITY; continue; }
        double ld = 0.0;
        for (int d = 0; d < D; d++) ld += log(var[(size_t)j * D + d]);
        for (int t = 0; t < T; t++) {
            double q = 0.0;
            for (int d = 0; d < D; d++) {
                double df = X[(size_t)t * D + d] - mean[(size_t)j * D + d];
                q += df * df / var[(size_t)j * D + d];
            }
            E[(size_t)t * S + j] = -0.5 * (D * ORC_LOG2PI + ld + q);
        }
    }
}

/* custom_hmm.py:176-211 */
void orc_forward(const double *E, int T, int S, const double *A, double *alpha, double *scale_out) {
    const double NINF = -INFINITY;
    for (size_t i = 0; i < (size_t)T * S; i++) alpha[i] = NINF;
    alpha[0] = 0.0;
    alpha[1] = log(A[0 * S + 1]) + E[1];
    for (int t = 1; t < T; t++) {
        const double *ap = alpha + (size_t)(t - 1) * S;
        double *ac = alpha + (size_t)t * S;
        ac[0] = NINF;
        for (int j = 1; j < S; j++) {
            if (j == 1)
                ac[j] = orc_logaddexp(ap[0] + log(A[0 * S + 1]), ap[1] + log(A[1 * S + 1])) + E[(size_t)t * S + j];
            else if (j < S - 1)
                ac[j] = orc_logaddexp(ap[j - 1] + log(A[(j - 1) * S + j]), ap[j] + log(A[j * S + j])) + E[(size_t)t * S + j];
            else
                ac[j] = ap[j - 1] + log(A[(j - 1) * S + j]);
        }
    }
    /* np.max propagates NaN */
    double mx = alpha[0];
    int has_nan = 0;
    for (size_t i = 0; i < (size_t)T * S; i++) {
        if (alpha[i] != alpha[i]) has_nan = 1;
        else if (alpha[i] > mx) mx = alpha[i];
    }
    if (has_nan) mx = NAN;
    for (size_t i = 0; i < (size_t)T * S; i++) alpha[i] -= mx;
    *scale_out = mx;
}

/* custom_hmm.py:213-246 */
void orc_backward(const double *E, int T, int S, const double *A, double scale, double *beta) {
    const double NINF = -INFINITY;
    for (size_t i = 0; i < (size_t)T * S; i++) beta[i] = NINF;
    beta[(size_t)(T - 1) * S + S - 1] = 0.0;
    for (int t = T - 2; t >= 0; t--) {
        const double *bn = beta + (size_t)(t + 1) * S;
        const double *en = E + (size_t)(t + 1) * S;
        double *bc = beta + (size_t)t * S;
        for (int i = 0; i < S - 1; i++) {
            if (i == 0)
                bc[i] = log(A[0 * S + 1]) + en[1] + bn[1];
            else if (i < S - 2)
                bc[i] = orc_logaddexp(log(A[i * S + i]) + en[i] + bn[i],
                                      log(A[i * S + i + 1]) + en[i + 1] + bn[i + 1]);
            else
                bc[i] = orc_logaddexp(log(A[i * S + i]) + en[i] + bn[i],
                                      log(A[i * S + i + 1]) + bn[i + 1]);
        }
    }
    for (size_t i = 0; i < (size_t)(T - 1) * S; i++) beta[i] -= scale;
}

/* custom_hmm.py:248-257 */
void orc_gamma(const double *alpha, const double *beta, int T, int S, double *gamma) {
    double *lg = (double *)malloc(sizeof(double) * S);
    for (int t = 0; t < T; t++) {
        for (int j = 0; j < S; j++) lg[j] = alpha[(size_t)t * S + j] + beta[(size_t)t * S + j];
        double norm = orc_logaddexp_reduce(lg, S);
        for (int j = 0; j < S; j++) gamma[(size_t)t * S + j] = exp(lg[j] - norm);
    }
    free(lg);
}

/* custom_hmm.py:259-322; xi is (T-1) x S x S */
void orc_xi(const double *alpha, const double *beta, const double *E, int T, int S,
            const double *A, double *xi) {
    memset(xi, 0, sizeof(double) * (size_t)(T - 1) * S * S);
    double ll = orc_logaddexp_reduce(alpha + (size_t)(T - 1) * S, S);
    for (int t = 0; t < T - 1; t++) {
        double *x = xi + (size_t)t * S * S;
        const double *a = alpha + (size_t)t * S;
        const double *bn = beta + (size_t)(t + 1) * S;
        const double *en = E + (size_t)(t + 1) * S;
        x[0 * S + 1] = exp(a[0] + log(A[0 * S + 1]) + en[1] + bn[1] - ll);
        for (int i = 1; i < S - 1; i++) {
            if (A[i * S + i] > 0)
                x[i * S + i] = exp(a[i] + log(A[i * S + i]) + en[i] + bn[i] - ll);
            if (i < S - 2)
                x[i * S + i + 1] = exp(a[i] + log(A[i * S + i + 1]) + en[i + 1] + bn[i + 1] - ll);
        }
        x[(S - 2) * S + (S - 1)] = exp(a[S - 2] + log(A[(S - 2) * S + S - 1]) + en[S - 1] + bn[S - 1] - ll);
        x[(S - 1) * S + (S - 1)] = exp(a[S - 1] + log(A[(S - 1) * S + S - 1]) + en[S - 1] + bn[S - 1] - ll);
        double sum = 0.0;
        for (int k = 0; k < S * S; k++) sum += x[k];
        if (sum > 0)
            for (int k = 0; k < S * S; k++) x[k] /= sum;
    }
}

/* custom_hmm.py:351-364 */
void orc_update_A(const double *agg_xi, const double *agg_gamma, int S, double *A) {
    A[0 * S + 1] = 1.0;
    for (int i = 1; i < S - 1; i++)
        if (agg_gamma[i] > 0) {
            A[i * S + i] = agg_xi[i * S + i] / agg_gamma[i];
            A[i * S + i + 1] = 1.0 - A[i * S + i];
        }
    A[(S - 1) * S + S - 1] = 1.0;
}

/* custom_hmm.py:366-400: two-pass full-covariance M-step.  gamma is the
 * concatenation of the per-utterance (T_u x S) matrices. global_cov is D x D. */
void orc_update_B(const double *X, const int64_t *offsets, int B, int D, int S,
                  const double *gamma, const double *global_cov, double var_floor_factor,
                  double *mean, double *cov) {
    int64_t total = offsets[B];
    double *occ = (double *)calloc(S, sizeof(double));
    memset(mean, 0, sizeof(double) * S * D);
    memset(cov, 0, sizeof(double) * (size_t)S * D * D);
    for (int64_t t = 0; t < total; t++)
        for (int j = 1; j < S - 1; j++) {
            double g = gamma[t * S + j];
            occ[j] += g;
            for (int d = 0; d < D; d++) mean[(size_t)j * D + d] += g * X[t * D + d];
        }
    for (int j = 1; j < S - 1; j++)
        if (occ[j] > 0)
            for (int d = 0; d < D; d++) mean[(size_t)j * D + d] /= occ[j];
    double *df = (double *)malloc(sizeof(double) * D);
    for (int64_t t = 0; t < total; t++)
        for (int j = 1; j < S - 1; j++) {
            double g = gamma[t * S + j];
            for (int d = 0; d < D; d++) df[d] = X[t * D + d] - mean[(size_t)j * D + d];
            double *cj = cov + (size_t)j * D * D;
            for (int a = 0; a < D; a++)
                for (int b = 0; b < D; b++) cj[a * D + b] += g * (df[a] * df[b]);
        }
    free(df);
    double tr = 0.0;
    for (int d = 0; d < D; d++) tr += global_cov[d * D + d];
    double floor_v = var_floor_factor * (tr / D);
    for (int j = 1; j < S - 1; j++)
        if (occ[j] > 0) {
            double *cj = cov + (size_t)j * D * D;
            for (int k = 0; k < D * D; k++) cj[k] /= occ[j];
            for (int a = 0; a < D; a++)
                for (int b = a + 1; b < D; b++) {
                    double m = (cj[a * D + b] + cj[b * D + a]) / 2;
                    cj[a * D + b] = m; cj[b * D + a] = m;
                }
            for (int a = 0; a < D; a++) {
                cj[a * D + a] = (cj[a * D + a] + cj[a * D + a]) / 2;
                if (!(cj[a * D + a] >= floor_v)) /* np.maximum propagates NaN: keep NaN */
                    if (cj[a * D + a] == cj[a * D + a]) cj[a * D + a] = floor_v;
            }
        }
    free(occ);
}

/* custom_hmm.py:462-514 given an emission matrix (T_eff rows are walked).
 * Returns V[T_eff-1, S-1]; path has T_eff entries.                           */
double orc_decode(const double *E, int T_eff, int S, const double *A, int32_t *path) {
    const int N = S - 2;
    const double NINF = -INFINITY;
    double *V = (double *)malloc(sizeof(double) * (size_t)T_eff * S);
    int32_t *bp = (int32_t *)calloc((size_t)T_eff * S, sizeof(int32_t));
    for (size_t i = 0; i < (size_t)T_eff * S; i++) V[i] = NINF;
    V[0] = 0.0;
    V[1] = log(A[0 * S + 1]) + E[1];
    for (int t = 1; t < T_eff; t++) {
        for (int j = 1; j < S; j++) {
            int cand[2], nc = 0;
            if (j == 1) { cand[nc++] = 1; if (t == 1) cand[nc++] = 0; }
            else if (j == S - 1) { if (t >= N) { cand[nc++] = j - 1; cand[nc++] = j; } else continue; }
            else { cand[nc++] = j - 1; cand[nc++] = j; }
            double best = NINF; int best_prev = -1;
            for (int c = 0; c < nc; c++) {
                int i = cand[c];
                double score = V[(size_t)(t - 1) * S + i] + log(A[i * S + j]);
                if (score > best) { best = score; best_prev = i; }
            }
            if (best_prev >= 0) {
                V[(size_t)t * S + j] = (j != S - 1) ? best + E[(size_t)t * S + j] : best;
                bp[(size_t)t * S + j] = best_prev;
            }
        }
    }
    int cur = S - 1;
    for (int t = T_eff - 1; t >= 0; t--) { path[t] = cur; cur = bp[(size_t)t * S + cur]; }
    double score = V[(size_t)(T_eff - 1) * S + S - 1];
    free(bp); free(V);
    return score;
}

/* One E-step over a batch with ONE model (custom_hmm.py:417-439): accumulates
 * agg_gamma[S] (sum of gamma[:-1]), agg_xi[S*S], per-utterance LL (D6
 * definition) and writes the concatenated gamma (needed by update_B).
 * emission_mode: 0 = sapr Gram/full-cov (cov is S x D x D), 1 = diag (cov is
 * S x D variances).                                                          */
void orc_estep_model(const double *X, const int64_t *offsets, int B, int D, int S,
                     int emission_mode, const double *A, const double *mean, const double *cov,
                     double *agg_gamma, double *agg_xi, double *loglik, double *gamma_out) {
    memset(agg_gamma, 0, sizeof(double) * S);
    memset(agg_xi, 0, sizeof(double) * S * S);
    for (int u = 0; u < B; u++) {
        int T = (int)(offsets[u + 1] - offsets[u]);
        const double *Xu = X + (size_t)offsets[u] * D;
        double *E = (double *)malloc(sizeof(double) * (size_t)T * S);
        double *al = (double *)malloc(sizeof(double) * (size_t)T * S);
        double *be = (double *)malloc(sizeof(double) * (size_t)T * S);
        double *xi = (double *)malloc(sizeof(double) * (size_t)(T > 1 ? T - 1 : 1) * S * S);
        double *ga = gamma_out + (size_t)offsets[u] * S;
        double scale;
        if (emission_mode == 0) orc_emission_sapr(Xu, T, D, S, mean, cov, E);
        else orc_emission_diag(Xu, T, D, S, mean, cov, 0, E);
        orc_forward(E, T, S, A, al, &scale);
        orc_backward(E, T, S, A, scale, be);
        orc_gamma(al, be, T, S, ga);
        orc_xi(al, be, E, T, S, A, xi);
        for (int t = 0; t < T - 1; t++)
            for (int j = 0; j < S; j++) agg_gamma[j] += ga[(size_t)t * S + j];
        for (int t = 0; t < T - 1; t++)
            for (int k = 0; k < S * S; k++) agg_xi[k] += xi[(size_t)t * S * S + k];
        loglik[u] = orc_logaddexp_reduce(al + (size_t)(T - 1) * S, S);
        free(xi); free(be); free(al); free(E);
    }
}

/* custom_hmm.py:402-460.  Parameters are updated in place; history[max_iter];
 * returns the number of history entries written.  emission_mode 0 reproduces
 * the reference as written; 1 is the Rung-1 ladder (true diagonal emission in
 * compute_emission_matrix, everything else untouched: update_B still produces
 * FULL covariances, of which the emission then reads only the diagonal).     */
int orc_baum_welch(const double *X, const int64_t *offsets, int B, int D, int N,
                   int emission_mode, int max_iter, double tol, double var_floor_factor,
                   const double *global_cov, double *A, double *mean, double *cov, double *history) {
    const int S = N + 2;
    int64_t total = offsets[B];
    double prev = -INFINITY;
    int n = 0;
    double *agg_gamma = (double *)malloc(sizeof(double) * S);
    double *agg_xi = (double *)malloc(sizeof(double) * S * S);
    double *ll = (double *)malloc(sizeof(double) * B);
    double *gamma = (double *)malloc(sizeof(double) * (size_t)total * S);
    double *var = (double *)malloc(sizeof(double) * S * D);
    for (int it = 0; it < max_iter; it++) {
        const double *covarg = cov;
        if (emission_mode == 1) {
            for (int j = 0; j < S; j++)
                for (int d = 0; d < D; d++) var[j * D + d] = cov[((size_t)j * D + d) * D + d];
            covarg = var;
        }
        orc_estep_model(X, offsets, B, D, S, emission_mode, A, mean, covarg, agg_gamma, agg_xi, ll, gamma);
        double tot = 0.0;
        for (int u = 0; u < B; u++) tot += ll[u];
        history[n++] = tot;
        if (fabs(tot - prev) < tol) break;
        prev = tot;
        orc_update_A(agg_xi, agg_gamma, S, A);
        orc_update_B(X, offsets, B, D, S, gamma, global_cov, var_floor_factor, mean, cov);
    }
    free(var); free(gamma); free(ll); free(agg_xi); free(agg_gamma);
    return n;
}

/* ------------------------------------------------------------------------- */
/* Batched, multi-threaded legs used by the parity tests at moderate sizes and
 * by bench.py's cpu_baseline / --impl reference.  Diagonal
 * emission + sapr topology (the cfg-2 / cfg-3 primary runs).                  */

/* decoder.py:35-49 over a batch: every utterance against all M models;
 * strict '>' keeps the first best model.  mean/var: M x S x D, A: M x S x S.
 * first_frames > 0 walks only that many frames.  Outputs:
 * best_word[B], best_score[B], scores[B*M] (nullable), best_path[sum T_eff]
 * (nullable, at the utterance's frame offset, T_eff entries used).           */
void orc_viterbi_batch(const double *X, const int64_t *offsets, int B, int D, int S, int M,
                       const double *A, const double *mean, const double *var, int first_frames,
                       int32_t *best_word, double *best_score, double *scores, int32_t *best_path,
                       int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int u = 0; u < B; u++) {
        int T = (int)(offsets[u + 1] - offsets[u]);
        int Te = (first_frames > 0 && first_frames < T) ? first_frames : T;
        const double *Xu = X + (size_t)offsets[u] * D;
        double *E = (double *)malloc(sizeof(double) * (size_t)T * S);
        int32_t *p = (int32_t *)malloc(sizeof(int32_t) * T);
        double bs = -INFINITY; int bw = -1;
        for (int m = 0; m < M; m++) {
            orc_emission_diag(Xu, T, D, S, mean + (size_t)m * S * D, var + (size_t)m * S * D, 0, E);
            double sc = orc_decode(E, Te, S, A + (size_t)m * S * S, p);
            if (scores) scores[(size_t)u * M + m] = sc;
            if (sc > bs) {
                bs = sc; bw = m;
                if (best_path) memcpy(best_path + offsets[u], p, sizeof(int32_t) * Te);
            }
        }
        best_word[u] = bw; best_score[u] = bs;
        free(p); free(E);
    }
}

/* E-step of custom_hmm.py:417-439 + the sums update_B consumes, batched over
 * utterances that each belong to model_of_utt[u]; diagonal statistics around a
 * per-state pivot (pivot = the model's current mean) exactly as the CUDA path
 * accumulates them.  stats layout per model (stride = 3*S + 2*S*D doubles):
 *   [0,S)        G   = sum_u sum_{t<=T-2} gamma_t(j)
 *   [S,2S)       Xs  = sum_u sum_t xi_t(j,j)
 *   [2S,3S)      occ = sum_u sum_t gamma_t(j)          (all T rows)
 *   [3S,3S+SD)   sum gamma (x - pivot)
 *   [3S+SD, ..)  sum gamma (x - pivot)^2
 * loglik[B] uses the D6 definition.                                          */
void orc_estep_batch(const double *X, const int64_t *offsets, int B, int D, int S, int M,
                     const int32_t *model_of_utt, const double *A, const double *mean,
                     const double *var, double *stats, double *loglik, int nthreads) {
    const size_t stride = (size_t)3 * S + (size_t)2 * S * D;
    memset(stats, 0, sizeof(double) * stride * M);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *loc = (double *)calloc(stride * M, sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int u = 0; u < B; u++) {
            int T = (int)(offsets[u + 1] - offsets[u]);
            int m = model_of_utt[u];
            const double *Xu = X + (size_t)offsets[u] * D;
            const double *Am = A + (size_t)m * S * S;
            const double *mu = mean + (size_t)m * S * D;
            const double *vr = var + (size_t)m * S * D;
            double *E = (double *)malloc(sizeof(double) * (size_t)T * S);
            double *al = (double *)malloc(sizeof(double) * (size_t)T * S);
            double *be = (double *)malloc(sizeof(double) * (size_t)T * S);
            double *ga = (double *)malloc(sizeof(double) * (size_t)T * S);
            double *xi = (double *)malloc(sizeof(double) * (size_t)(T > 1 ? T - 1 : 1) * S * S);
            double scale;
            orc_emission_diag(Xu, T, D, S, mu, vr, 0, E);
            orc_forward(E, T, S, Am, al, &scale);
            orc_backward(E, T, S, Am, scale, be);
            orc_gamma(al, be, T, S, ga);
            orc_xi(al, be, E, T, S, Am, xi);
            double *st = loc + stride * m;
            for (int t = 0; t < T; t++)
                for (int j = 1; j < S - 1; j++) {
                    double g = ga[(size_t)t * S + j];
                    if (t < T - 1) st[j] += g;
                    st[2 * S + j] += g;
                    for (int d = 0; d < D; d++) {
                        double df = Xu[(size_t)t * D + d] - mu[(size_t)j * D + d];
                        st[3 * S + (size_t)j * D + d] += g * df;
                        st[3 * S + (size_t)S * D + (size_t)j * D + d] += g * df * df;
                    }
                }
            for (int t = 0; t < T - 1; t++)
                for (int j = 1; j < S - 1; j++) st[S + j] += xi[((size_t)t * S + j) * S + j];
            loglik[u] = orc_logaddexp_reduce(al + (size_t)(T - 1) * S, S);
            free(xi); free(ga); free(be); free(al); free(E);
        }
#pragma omp critical
        for (size_t k = 0; k < stride * M; k++) stats[k] += loc[k];
        free(loc);
    }
}

/* ------------------------------------------------------------------------- */
/* hmmlearn 0.3.3 GaussianHMM(covariance_type="diag", implementation="log")
 * semantics -- PARITY UNPINNED (package absent).  S
 * all-emitting states, dense S x S transmat, startprob[S].  logA / logpi are
 * element-wise logs (log 0 = -inf).                                          */
static double orc_lse(const double *v, int n) {
    double m = -INFINITY;
    for (int i = 0; i < n; i++) if (v[i] > m) m = v[i];
    if (!(m > -INFINITY)) return -INFINITY;
    double s = 0.0;
    for (int i = 0; i < n; i++) s += exp(v[i] - m);
    return log(s) + m;
}

double orc_hl_forward(const double *lf, int T, int S, const double *logpi, const double *logA, double *fwd) {
    double *w = (double *)malloc(sizeof(double) * S);
    for (int i = 0; i < S; i++) fwd[i] = logpi[i] + lf[i];
    for (int t = 1; t < T; t++)
        for (int j = 0; j < S; j++) {
            for (int i = 0; i < S; i++) w[i] = fwd[(size_t)(t - 1) * S + i] + logA[i * S + j];
            fwd[(size_t)t * S + j] = orc_lse(w, S) + lf[(size_t)t * S + j];
        }
    double lp = orc_lse(fwd + (size_t)(T - 1) * S, S);
    free(w);
    return lp;
}

void orc_hl_backward(const double *lf, int T, int S, const double *logA, double *bwd) {
    double *w = (double *)malloc(sizeof(double) * S);
    for (int i = 0; i < S; i++) bwd[(size_t)(T - 1) * S + i] = 0.0;
    for (int t = T - 2; t >= 0; t--)
        for (int i = 0; i < S; i++) {
            for (int j = 0; j < S; j++)
                w[j] = logA[i * S + j] + lf[(size_t)(t + 1) * S + j] + bwd[(size_t)(t + 1) * S + j];
            bwd[(size_t)t * S + i] = orc_lse(w, S);
        }
    free(w);
}

double orc_hl_viterbi(const double *lf, int T, int S, const double *logpi, const double *logA, int32_t *path) {
    double *dl = (double *)malloc(sizeof(double) * (size_t)T * S);
    for (int i = 0; i < S; i++) dl[i] = logpi[i] + lf[i];
    for (int t = 1; t < T; t++)
        for (int j = 0; j < S; j++) {
            double m = -INFINITY;
            for (int i = 0; i < S; i++) {
                double v = dl[(size_t)(t - 1) * S + i] + logA[i * S + j];
                if (v > m) m = v;
            }
            dl[(size_t)t * S + j] = m + lf[(size_t)t * S + j];
        }
    int cur = 0; double best = dl[(size_t)(T - 1) * S + 0];
    for (int i = 1; i < S; i++) if (dl[(size_t)(T - 1) * S + i] > best) { best = dl[(size_t)(T - 1) * S + i]; cur = i; }
    path[T - 1] = cur;
    for (int t = T - 2; t >= 0; t--) {
        int arg = 0; double m = dl[(size_t)t * S] + logA[0 * S + cur];
        for (int i = 1; i < S; i++) {
            double v = dl[(size_t)t * S + i] + logA[i * S + cur];
            if (v > m) { m = v; arg = i; }
        }
        cur = arg; path[t] = cur;
    }
    free(dl);
    return best;
}

/* One hmmlearn E-step over a batch (one model): accumulates start[S],
 * trans[S*S], post[S], obs[S*D], obs2[S*D]; returns total log_prob.          */
double orc_hl_estep(const double *X, const int64_t *offsets, int B, int D, int S,
                    const double *startprob, const double *transmat, const double *mean,
                    const double *var, double *start, double *trans, double *post,
                    double *obs, double *obs2, double *loglik) {
    double *logA = (double *)malloc(sizeof(double) * S * S);
    double *logpi = (double *)malloc(sizeof(double) * S);
    for (int i = 0; i < S * S; i++) logA[i] = log(transmat[i]);
    for (int i = 0; i < S; i++) logpi[i] = log(startprob[i]);
    memset(start, 0, sizeof(double) * S); memset(trans, 0, sizeof(double) * S * S);
    memset(post, 0, sizeof(double) * S); memset(obs, 0, sizeof(double) * S * D);
    memset(obs2, 0, sizeof(double) * S * D);
    double total = 0.0;
    double *w = (double *)malloc(sizeof(double) * S);
    for (int u = 0; u < B; u++) {
        int T = (int)(offsets[u + 1] - offsets[u]);
        const double *Xu = X + (size_t)offsets[u] * D;
        double *lf = (double *)malloc(sizeof(double) * (size_t)T * S);
        double *fw = (double *)malloc(sizeof(double) * (size_t)T * S);
        double *bw = (double *)malloc(sizeof(double) * (size_t)T * S);
        orc_emission_diag(Xu, T, D, S, mean, var, 1, lf);
        double lp = orc_hl_forward(lf, T, S, logpi, logA, fw);
        orc_hl_backward(lf, T, S, logA, bw);
        total += lp;
        if (loglik) loglik[u] = lp;
        for (int t = 0; t < T; t++) {
            for (int j = 0; j < S; j++) w[j] = fw[(size_t)t * S + j] + bw[(size_t)t * S + j];
            double nrm = orc_lse(w, S);
            for (int j = 0; j < S; j++) {
                double g = exp(w[j] - nrm);
                if (t == 0) start[j] += g;
                post[j] += g;
                for (int d = 0; d < D; d++) {
                    double x = Xu[(size_t)t * D + d];
                    obs[(size_t)j * D + d] += g * x;
                    obs2[(size_t)j * D + d] += g * x * x;
                }
            }
        }
        if (T > 1)
            for (int i = 0; i < S; i++)
                for (int j = 0; j < S; j++) {
                    if (!(logA[i * S + j] > -INFINITY)) continue;
                    double m = -INFINITY;
                    for (int t = 0; t < T - 1; t++) {
                        double v = fw[(size_t)t * S + i] + logA[i * S + j] + lf[(size_t)(t + 1) * S + j] + bw[(size_t)(t + 1) * S + j] - lp;
                        if (v > m) m = v;
                    }
                    if (!(m > -INFINITY)) continue;
                    double s = 0.0;
                    for (int t = 0; t < T - 1; t++)
                        s += exp(fw[(size_t)t * S + i] + logA[i * S + j] + lf[(size_t)(t + 1) * S + j] + bw[(size_t)(t + 1) * S + j] - lp - m);
                    trans[i * S + j] += exp(log(s) + m);
                }
        free(bw); free(fw); free(lf);
    }
    free(w); free(logpi); free(logA);
    return total;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
