"""oracle/mel_np.py -- TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's log-mel front-end.

Follows /root/reference/src/qwen2-whisper.cpp:
  padding            log_mel_spectrogram            :2594-2606  (reflect 200 front, 30 s + 200 zeros back)
  n_len, n_len_org                                  :2611-2613
  Hann (periodic, cosf)  whisper_global_cache       :2428-2436
  frame -> FFT -> |X|^2  worker_thread              :2522-2542  (the reference's FFT is F32; restated in F64)
  filterbank in double, log10(max(., 1e-10))        :2545-2561
  frames past the signal = log10(1e-10)             :2566-2571
  global max, clamp max-8, (x+4)/4                  :2634-2649
Pinned against the reference itself by tests/test_oracle_cpu.py (max-abs <= 2e-5 on noisy and low-noise inputs).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE, N_FFT, HOP = 16000, 400, 160


def hann_periodic() -> np.ndarray:
    i = np.arange(N_FFT)
    c = np.cos(((2.0 * np.pi * i) / N_FFT).astype(np.float32).astype(np.float64)).astype(np.float32)   # cosf(float(theta))
    return (0.5 * (1.0 - c.astype(np.float64))).astype(np.float32)


def pad_signal(pcm: np.ndarray) -> np.ndarray:
    pcm = np.asarray(pcm, dtype=np.float32)
    n = pcm.size
    x = np.zeros(n + SAMPLE_RATE * 30 + 2 * (N_FFT // 2), dtype=np.float32)
    x[N_FFT // 2:N_FFT // 2 + n] = pcm
    x[:N_FFT // 2] = pcm[1:1 + N_FFT // 2][::-1]
    return x


def mel_dims(n_samples: int):
    n_len = (n_samples + SAMPLE_RATE * 30 + N_FFT - N_FFT) // HOP
    n_len_org = 1 + (n_samples + N_FFT // 2 - N_FFT) // HOP
    return n_len, n_len_org


def log_mel_unnormalised(pcm: np.ndarray, filters: np.ndarray, n_frames: int | None = None) -> np.ndarray:
    """float64 [n_mel, n_len] of log10(max(mel energy, 1e-10)) before the clamp"""
    x = pad_signal(pcm)
    n = np.asarray(pcm).size
    n_len, _ = mel_dims(n)
    if n_frames is not None:
        n_len = min(n_len, n_frames)
    n_arg = n + N_FFT // 2                               # the worker's n_samples argument (:2621)
    n_calc = min(n_arg // HOP + 1, n_len)                # frames that run the FFT (:2522)
    hann = hann_periodic().astype(np.float64)
    out = np.full((filters.shape[0], n_len), np.log10(1e-10), dtype=np.float64)
    xz = x.astype(np.float64).copy()
    xz[n_arg:] = 0.0                                     # samples past n_arg are never read (:2526-2533)
    idx = (np.arange(n_calc) * HOP)[:, None] + np.arange(N_FFT)[None, :]
    frames = xz[idx] * hann[None, :]
    spec = np.fft.rfft(frames, axis=1)
    power = (spec.real ** 2 + spec.imag ** 2)
    mel = power @ filters.astype(np.float64).T           # [n_calc, n_mel]
    out[:, :n_calc] = np.log10(np.maximum(mel, 1e-10)).T
    return out


def log_mel_spectrogram(pcm: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """float32 [n_mel, n_len], the reference's whisper_mel.data"""
    m = log_mel_unnormalised(pcm, filters).astype(np.float32).astype(np.float64)   # mel.data is float (:2561)
    mmax = m.max() - 8.0
    m = np.maximum(m, mmax)
    return ((m + 4.0) / 4.0).astype(np.float32)


def window(mel: np.ndarray, offset: int, n_ctx: int) -> np.ndarray:
    """the [n_mel, 2*n_ctx] encoder input: frames [offset, offset+2*n_ctx), zero-filled past n_len (:2274-2283)"""
    n_mel, n_len = mel.shape
    out = np.zeros((n_mel, 2 * n_ctx), dtype=np.float32)
    i0, i1 = min(offset, n_len), min(offset + 2 * n_ctx, n_len)
    out[:, :i1 - i0] = mel[:, i0:i1]
    return out
