"""Drop-in for assignment2/mfcc_extract.py: same function names and on-disk format ((13, T) float32
``.npy``, word = last ``_`` token of the file name), MFCCs computed by the fused CUDA kernels of
csrc/mfcc.cu instead of librosa.

Out of scope: audio *decoding and resampling* (librosa.load -> audioread/soxr).
``extract_mfcc`` therefore reads PCM ``.wav`` (stdlib ``wave``) or ``.npy`` sample arrays and uses the
file's own sample rate; the reference resamples everything to 22 050 Hz first.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import wave

import numpy as np

from . import _lib
from ._lib import MfccParams, ptr

logging.basicConfig(level=logging.DEBUG)   # mfcc_extract.py:6


def librosa_params(sr: int = 22050) -> MfccParams:
    """What librosa.feature.mfcc computes for the call in mfcc_extract.py:13-23."""
    return MfccParams(sample_rate=sr, n_fft=2048, win_length=int(0.03 * sr), hop_length=int(0.01 * sr), n_mels=128,
                      n_mfcc=13, center=1, mel_slaney=1, log_db=1, top_db=80.0, preemph=0.0, fmin=0.0, fmax=sr / 2.0)


def cfg5_params() -> MfccParams:
    """BASELINE config 5: 16 kHz, 25 ms / 10 ms Hamming, 512-pt FFT, 26 mel, DCT-13, pre-emphasis 0.97."""
    return MfccParams(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=26, n_mfcc=13, center=0,
                      mel_slaney=0, log_db=0, top_db=0.0, preemph=0.97, fmin=0.0, fmax=8000.0)


def mfcc_batch(audio, sample_offsets, params: MfccParams, ld_out=None, ctx=None):
    """audio: float32 CUDA tensor [sum_samples]; sample_offsets: int64 host array [B+1].
    Returns (feats [sum_frames, ld_out] float32 CUDA, frame_offsets int64 host [B+1])."""
    import torch

    ctx = ctx or _lib.default_context()
    so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    B = len(so) - 1
    fo = np.zeros(B + 1, dtype=np.int64)
    for u in range(B):
        fo[u + 1] = fo[u] + ctx.lib.sapr_mfcc_num_frames(C.byref(params), int(so[u + 1] - so[u]))
    ld = ld_out or ((params.n_mfcc + 3) // 4 * 4)
    feats = torch.zeros((int(fo[-1]), ld), dtype=torch.float32, device=audio.device)
    fo2 = np.zeros(B + 1, dtype=np.int64)
    ctx.check(ctx.lib.sapr_mfcc(ctx.h, C.byref(params), ptr(audio), ptr(so), B, ptr(feats), ld, ptr(fo2)))
    assert np.array_equal(fo, fo2)
    return feats, fo2


def mfcc_from_samples(y: np.ndarray, sr: int, params: MfccParams = None) -> np.ndarray:
    """(samples,) -> (n_mfcc, frames) float32, the array extract_mfcc returns."""
    import torch

    p = params or librosa_params(sr)
    a = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32), device="cuda")
    feats, fo = mfcc_batch(a, np.array([0, len(y)], dtype=np.int64), p)
    return np.ascontiguousarray(feats[:, :p.n_mfcc].cpu().numpy().T)


def _load_audio(path: str):
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32), 22050
    with wave.open(path, "rb") as w:
        sr, n, width, ch = w.getframerate(), w.getnframes(), w.getsampwidth(), w.getnchannels()
        raw = w.readframes(n)
    if width != 2:
        raise ValueError("only 16-bit PCM wav is supported")
    y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    if ch > 1:
        y = y.reshape(-1, ch).mean(axis=1)
    return y, sr


def extract_mfcc(audio_path: str) -> np.ndarray:
    try:
        y, sr = _load_audio(audio_path)
        return mfcc_from_samples(y, sr)
    except Exception as e:
        logging.error(f"Error processing {audio_path}: {str(e)}")
        raise


def extract_mfccs(input_folder: str, output_folder: str, ext: str = ".wav") -> str:
    logging.debug(f"Extracting MFCCs from {input_folder} to {output_folder}...")
    os.makedirs(output_folder, exist_ok=True)
    processed_files = 0
    for file in os.listdir(input_folder):
        if file.endswith(ext):
            try:
                mfcc = extract_mfcc(os.path.join(input_folder, file))
                np.save(os.path.join(output_folder, file[: -len(ext)] + ".npy"), mfcc)
                processed_files += 1
            except Exception as e:
                logging.error(f"Failed to process {file}: {str(e)}")
                continue
    logging.info(f"Completed processing {processed_files} files")
    return output_folder


def load_mfcc(file_path: str) -> np.ndarray:
    try:
        return np.load(file_path)
    except Exception as e:
        logging.error(f"Failed to load MFCC from {file_path}: {str(e)}")
        raise


def load_mfccs(directory_path: str) -> list:
    feature_list = []
    for file_name in os.listdir(directory_path):
        if file_name.endswith(".npy"):
            feature_list.append(load_mfcc(os.path.join(directory_path, file_name)))
    return feature_list


def load_mfccs_by_word(directory_path: str, word: str) -> list:
    mfccs = []
    for file_name in os.listdir(directory_path):
        file_word = file_name.split("_")[-1].split(".")[0]
        if file_name.endswith(".npy") and file_word == word:
            mfccs.append(load_mfcc(os.path.join(directory_path, file_name)))
    return mfccs
