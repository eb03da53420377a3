// host.cu -- host-buffer entry point: the call a reference-side plugin makes when its features live
// in numpy arrays (decoder.py:51-72 walks a list of per-utterance arrays; here they arrive packed).
// Utterance chunks are streamed H2D on a copy stream while the previous chunk decodes, results are
// copied back per chunk: the PCIe transfer, not the kernel, bounds this path.
#include "common.cuh"

extern "C" int sapr_viterbi_host(sapr_ctx *ctx, sapr_models *m, const float *X_host, int ldx, const int64_t *offsets_host,
                                 int B, int precision, int first_frames, int chunk_utts, int32_t *best_word_host,
                                 double *best_score_host, uint8_t *best_path_host) {
    if (!ctx || !m || !X_host || !offsets_host || !best_word_host || !best_score_host) return SAPR_E_INVALID;
    if (B <= 0) return SAPR_OK;
    if (chunk_utts <= 0) chunk_utts = 8192;
    chunk_utts = std::min(chunk_utts, B);
    // per-chunk extents
    int64_t max_frames = 0;
    int max_T = 0;
    for (int u0 = 0; u0 < B; u0 += chunk_utts) {
        const int u1 = std::min(B, u0 + chunk_utts);
        max_frames = std::max(max_frames, offsets_host[u1] - offsets_host[u0]);
    }
    for (int u = 0; u < B; u++) {
        const int64_t T = offsets_host[u + 1] - offsets_host[u];
        if (T < 0 || T > INT32_MAX) SAPR_FAIL(ctx, SAPR_E_INVALID, "viterbi_host: bad offsets");
        if (first_frames > 0 && T < first_frames) SAPR_FAIL(ctx, SAPR_E_SHORT, "viterbi_host: utterance shorter than first_frames");
        max_T = std::max(max_T, (int)T);
    }
    if (max_T <= 0) SAPR_FAIL(ctx, SAPR_E_INVALID, "viterbi_host: empty utterances");
    int rc;
    const size_t xbytes = sizeof(float) * (size_t)max_frames * ldx;
    const size_t obytes = sizeof(int64_t) * (size_t)(chunk_utts + 1);
    const size_t rbytes = (size_t)chunk_utts * (sizeof(int32_t) + sizeof(double)) + (size_t)max_frames + 64;
    if ((rc = sapr_ws_reserve(ctx, 2, 2 * xbytes))) return rc;
    if ((rc = sapr_ws_reserve(ctx, 3, 2 * (obytes + rbytes) + 256))) return rc;
    if ((rc = sapr_pin_reserve(ctx, 0, 2 * obytes))) return rc;
    // warm the kernel's own workspaces before the pipeline starts (they may synchronise when they grow)
    float *dX[2] = {(float *)ctx->ws[2], (float *)((char *)ctx->ws[2] + xbytes)};
    char *aux = (char *)ctx->ws[3];
    int64_t *dOff[2];
    double *dScore[2];
    int32_t *dWord[2];
    uint8_t *dPath[2];
    for (int s = 0; s < 2; s++) {
        char *p = aux + (size_t)s * ((obytes + rbytes + 127) / 128 * 128);
        dOff[s] = (int64_t *)p; p += (obytes + 15) / 16 * 16;
        dScore[s] = (double *)p; p += sizeof(double) * chunk_utts;
        dWord[s] = (int32_t *)p; p += sizeof(int32_t) * chunk_utts;
        dPath[s] = (uint8_t *)p;
    }
    int64_t *hOff[2] = {(int64_t *)ctx->pin[0], (int64_t *)((char *)ctx->pin[0] + obytes)};
    cudaEvent_t ready[2] = {ctx->ev[0], ctx->ev[1]}, freed[2] = {ctx->ev[2], ctx->ev[3]}, hostoff[2] = {ctx->ev[4], ctx->ev[5]};
    int c = 0;
    for (int u0 = 0; u0 < B; u0 += chunk_utts, c++) {
        const int s = c & 1;
        const int nu = std::min(chunk_utts, B - u0);
        const int64_t f0 = offsets_host[u0], nf = offsets_host[u0 + nu] - f0;
        if (c >= 2) {
            SAPR_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, freed[s], 0));
            SAPR_CUDA(ctx, cudaEventSynchronize(hostoff[s]));   // pinned offsets slot no longer being read
        }
        for (int i = 0; i <= nu; i++) hOff[s][i] = offsets_host[u0 + i] - f0;
        SAPR_CUDA(ctx, cudaMemcpyAsync(dOff[s], hOff[s], sizeof(int64_t) * (nu + 1), cudaMemcpyHostToDevice, ctx->copy_stream));
        SAPR_CUDA(ctx, cudaEventRecord(hostoff[s], ctx->copy_stream));
        SAPR_CUDA(ctx, cudaMemcpyAsync(dX[s], X_host + (size_t)f0 * ldx, sizeof(float) * (size_t)nf * ldx,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
        SAPR_CUDA(ctx, cudaEventRecord(ready[s], ctx->copy_stream));
        SAPR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ready[s], 0));
        rc = sapr_viterbi(ctx, m, dX[s], ldx, dOff[s], nu, nf, max_T, nullptr, precision, first_frames, dWord[s], dScore[s],
                          nullptr, best_path_host ? dPath[s] : nullptr, nullptr);
        if (rc) return rc;
        SAPR_CUDA(ctx, cudaMemcpyAsync(best_word_host + u0, dWord[s], sizeof(int32_t) * nu, cudaMemcpyDeviceToHost, ctx->stream));
        SAPR_CUDA(ctx, cudaMemcpyAsync(best_score_host + u0, dScore[s], sizeof(double) * nu, cudaMemcpyDeviceToHost, ctx->stream));
        if (best_path_host)
            SAPR_CUDA(ctx, cudaMemcpyAsync(best_path_host + f0, dPath[s], (size_t)nf, cudaMemcpyDeviceToHost, ctx->stream));
        SAPR_CUDA(ctx, cudaEventRecord(freed[s], ctx->stream));
    }
    SAPR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SAPR_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return SAPR_OK;
}
