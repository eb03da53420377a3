"""float64 numpy restatement of librosa.feature.mfcc as called by assignment2/mfcc_extract.py:13-23 --
TEST INFRASTRUCTURE, NOT PRODUCT CODE.  librosa 0.10.2.post1 (assignment2/poetry.lock:679-680) is not
vendored in the reference and not installed here: PARITY UNPINNED (the reference's own tests assert only
the output shape, tests/test_mfcc_extract.py:31-34).  Follows SURVEY.md Appendix C; parameterised like
sapr_mfcc_params so BASELINE cfg 5 is covered too."""
import numpy as np


def _hz_to_mel(f, slaney):
    f = np.asarray(f, dtype=np.float64)
    if not slaney:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m, slaney):
    m = np.asarray(m, dtype=np.float64)
    if not slaney:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax, slaney):
    pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin, slaney), _hz_to_mel(fmax, slaney), n_mels + 2), slaney)
    fft_f = np.arange(n_fft // 2 + 1) * sr / n_fft
    W = np.zeros((n_mels, n_fft // 2 + 1))
    for m in range(n_mels):
        lower = (fft_f - pts[m]) / (pts[m + 1] - pts[m])
        upper = (pts[m + 2] - fft_f) / (pts[m + 2] - pts[m + 1])
        W[m] = np.maximum(0, np.minimum(lower, upper))
        if slaney:
            W[m] *= 2.0 / (pts[m + 2] - pts[m])
    return W


def mfcc(y, sample_rate, n_fft, win_length, hop_length, n_mels, n_mfcc, center=1, mel_slaney=1, log_db=1,
         top_db=80.0, preemph=0.0, fmin=0.0, fmax=None):
    y = np.asarray(y, dtype=np.float64)
    fmax = fmax or sample_rate / 2
    if preemph:
        y = np.concatenate([y[:1], y[1:] - preemph * y[:-1]])
    win = np.zeros(n_fft)
    lpad = (n_fft - win_length) // 2
    win[lpad:lpad + win_length] = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(win_length) / win_length)
    if center:
        y = np.pad(y, n_fft // 2)
        n_frames = 1 + (len(y) - n_fft) // hop_length
    else:
        n_frames = 1 + (len(y) - n_fft) // hop_length if len(y) >= n_fft else 0
    frames = np.stack([y[i * hop_length:i * hop_length + n_fft] for i in range(n_frames)]) * win
    P = np.abs(np.fft.rfft(frames, axis=1)) ** 2
    S = P @ mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax, mel_slaney).T
    S = np.maximum(S, 1e-10)
    if log_db:
        L = 10.0 * np.log10(S)
        if top_db and top_db > 0:
            L = np.maximum(L, L.max() - top_db)
    else:
        L = np.log(S)
    n = np.arange(n_mels)
    k = np.arange(n_mfcc)[:, None]
    Dm = np.cos(np.pi * k * (2 * n + 1) / (2 * n_mels)) * np.sqrt(2.0 / n_mels)
    Dm[0] *= np.sqrt(0.5)
    return (L @ Dm.T).T      # (n_mfcc, frames)
