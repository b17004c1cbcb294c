"""Synthetic inputs of SURVEY 8(d): seeded PCM windows, random-init weights, the Slaney mel filterbank."""
from __future__ import annotations

import numpy as np

from . import modelfile as mfmod

FULL_HPARAMS = dict(n_vocab=51866, n_audio_ctx=1500, n_audio_state=1280, n_audio_head=20, n_audio_layer=32,
                    n_text_ctx=448, n_text_state=1280, n_text_head=20, n_text_layer=0, n_mels=128, ftype=1)
TINY_HPARAMS = dict(n_vocab=64, n_audio_ctx=100, n_audio_state=128, n_audio_head=2, n_audio_layer=2,
                    n_text_ctx=8, n_text_state=128, n_text_head=2, n_text_layer=0, n_mels=128, ftype=1)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    log_t = f >= 1000.0
    return np.where(log_t, 15.0 + np.log(np.maximum(f, 1e-9) / 1000.0) * (27.0 / np.log(6.4)), mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    log_t = m >= 15.0
    return np.where(log_t, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), f)


def slaney_mel_filters(n_mels: int = 128, n_fft: int = 400, sr: int = 16000) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular bank [n_mels, n_fft/2+1] -- what whisper's mel_filters.npz holds and
    models/convert-pt-to-ggml.py:268-281 writes into the file (== transformers WhisperFeatureExtractor.mel_filters.T)."""
    n_bins = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    hz = _mel_to_hz(mel_pts)
    fdiff = np.diff(hz)
    ramps = hz[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (hz[2:n_mels + 2] - hz[:n_mels]))[:, None]
    return w.astype(np.float32)


def synth_pcm(n: int = 480000, seed: int = 0, kind: str = "chirp") -> np.ndarray:
    """float32 PCM in [-1,1], round-tripped through int16 like read_wav (examples/common.cpp:723-728)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / 16000.0
    dur = max(n / 16000.0, 1e-9)
    if kind == "chirp":      # linear chirp 200 -> 500 Hz at 0.3 + N(0, 0.05^2)
        x = 0.3 * np.sin(2 * np.pi * (200.0 * t + 0.5 * (300.0 / dur) * t * t)) + rng.normal(0.0, 0.05, n)
    elif kind == "tones":    # low-noise: 440 Hz tone, silence, the same tone at 1e-3, silence
        x = np.zeros(n)
        q = n // 4
        x[:q] = 0.5 * np.sin(2 * np.pi * 440.0 * t[:q])
        x[2 * q:3 * q] = 1e-3 * np.sin(2 * np.pi * 440.0 * t[2 * q:3 * q])
    elif kind == "silence":
        x = np.zeros(n)
    elif kind == "noise":
        x = rng.normal(0.0, 0.2, n)
    else:
        raise ValueError(kind)
    i16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    return (i16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


_WEIGHT_CACHE: dict = {}   # the full-size set takes ~10 s to draw; tests and bench ask for the same (hparams, seed) many times


def synth_weights(hp: dict, seed: int = 1234) -> dict:
    """name -> float32 array (numpy shape): matrices N(0, 1/fan_in), biases / pos-emb N(0, 0.02^2), LN gamma 1 + N(0, 0.02^2).
    The returned arrays are shared between callers (cached): treat them as read-only."""
    key = (tuple(sorted(hp.items())), seed)
    if key in _WEIGHT_CACHE:
        return _WEIGHT_CACHE[key]
    out = _synth_weights(hp, seed)
    if len(_WEIGHT_CACHE) >= 2:
        _WEIGHT_CACHE.pop(next(iter(_WEIGHT_CACHE)))
    _WEIGHT_CACHE[key] = out
    return out


def _synth_weights(hp: dict, seed: int) -> dict:
    rng = np.random.default_rng(seed)
    out = {}
    for name, ne in mfmod.expected_shapes(hp).items():
        shape = tuple(reversed(ne))
        if name.endswith("layer_norm.weight"):
            w = 1.0 + 0.02 * rng.standard_normal(shape, dtype=np.float32)
        elif name.endswith(".bias") or name == "embed_positions.weight":
            w = 0.02 * rng.standard_normal(shape, dtype=np.float32)
        else:
            fan_in = int(np.prod(ne[:-1]))
            w = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / np.sqrt(fan_in))
        out[name] = w.astype(np.float32)
    return out


def synth_model(hp: dict, wtype: int, seed: int = 1234, filters: np.ndarray | None = None) -> mfmod.ModelFile:
    if filters is None:
        filters = slaney_mel_filters(hp["n_mels"])
    return mfmod.build_model(hp, filters, synth_weights(hp, seed), wtype)
