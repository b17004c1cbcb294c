"""oracle/refbind.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding to oracle/_ref/libq2wref_{v3,v4}.so: the UNMODIFIED reference (ggml CPU backend +
src/qwen2-whisper.cpp) compiled by oracle/Makefile through oracle/ref_harness.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
Nothing here reads /root/reference at run time: the .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def variant() -> str:
    fl = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"}
    return "v4" if v4 <= fl else "v3"


def lib_path() -> str:
    return os.path.join(REF_DIR, f"libq2wref_{variant()}.so")


def available() -> bool:
    return os.path.exists(lib_path())


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise FileNotFoundError(f"{p} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(p)
        vp, i = C.c_void_p, C.c_int
        L.q2wref_set_quiet.argtypes = [i]
        L.q2wref_init_from_file.restype = vp; L.q2wref_init_from_file.argtypes = [C.c_char_p]
        L.q2wref_init_from_buffer.restype = vp; L.q2wref_init_from_buffer.argtypes = [vp, C.c_size_t]
        L.q2wref_free.argtypes = [vp]
        L.q2wref_pcm_to_mel.restype = i; L.q2wref_pcm_to_mel.argtypes = [vp, vp, i, i]
        L.q2wref_set_mel.restype = i; L.q2wref_set_mel.argtypes = [vp, vp, i, i]
        L.q2wref_mel_dims.argtypes = [vp, C.POINTER(i)]
        L.q2wref_get_mel.argtypes = [vp, vp]
        L.q2wref_full.restype = i; L.q2wref_full.argtypes = [vp, vp, i, i, i]
        L.q2wref_embd_dims.restype = C.c_long; L.q2wref_embd_dims.argtypes = [vp, C.POINTER(i)]
        L.q2wref_get_embd.restype = i; L.q2wref_get_embd.argtypes = [vp, vp]
        L.q2wref_timings.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.q2wref_reset_timings.argtypes = [vp]
        L.q2wref_hparams.argtypes = [vp, C.POINTER(i)]
        L.q2wref_quantize.restype = C.c_size_t; L.q2wref_quantize.argtypes = [i, vp, vp, C.c_long, C.c_long]
        L.q2wref_dequantize.argtypes = [i, vp, vp, C.c_long]
        L.q2wref_row_size.restype = C.c_size_t; L.q2wref_row_size.argtypes = [i, C.c_long]
        L.q2wref_gelu.argtypes = [vp, vp, i]
        L.q2wref_set_quiet(1)
        _lib = L
    return _lib


class RefContext:
    """The reference's whisper_context on its ggml CPU backend."""

    def __init__(self, model: bytes | str):
        L = load()
        if isinstance(model, str):
            self._h = L.q2wref_init_from_file(model.encode())
            self._buf = None
        else:
            self._buf = (C.c_char * len(model)).from_buffer_copy(model)
            self._h = L.q2wref_init_from_buffer(C.cast(self._buf, C.c_void_p), len(model))
        if not self._h:
            raise RuntimeError("reference whisper_init_* returned NULL")

    def free(self):
        if self._h:
            load().q2wref_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def pcm_to_mel(self, pcm, n_threads=4) -> np.ndarray:
        p = np.ascontiguousarray(pcm, dtype=np.float32)
        rc = load().q2wref_pcm_to_mel(self._h, p.ctypes.data, p.size, n_threads)
        if rc != 0:
            raise RuntimeError(f"reference whisper_pcm_to_mel -> {rc}")
        return self.get_mel()

    def mel_dims(self):
        d = (C.c_int * 3)()
        load().q2wref_mel_dims(self._h, d)
        return d[0], d[1], d[2]

    def get_mel(self) -> np.ndarray:
        n_len, _, n_mel = self.mel_dims()
        out = np.empty((n_mel, n_len), dtype=np.float32)
        load().q2wref_get_mel(self._h, out.ctypes.data)
        return out

    def set_mel(self, mel: np.ndarray) -> int:
        m = np.ascontiguousarray(mel, dtype=np.float32)
        return load().q2wref_set_mel(self._h, m.ctypes.data, m.shape[1], m.shape[0])

    def full(self, pcm=None, n_threads=4, offset_ms=0) -> int:
        if pcm is None:
            return load().q2wref_full(self._h, None, 0, n_threads, offset_ms)
        p = np.ascontiguousarray(pcm, dtype=np.float32)
        return load().q2wref_full(self._h, p.ctypes.data, p.size, n_threads, offset_ms)

    def get_embeddings(self) -> np.ndarray:
        d = (C.c_int * 2)()
        n = load().q2wref_embd_dims(self._h, d)
        if n < 0:
            raise RuntimeError("reference has no embd_enc yet")
        out = np.empty((d[1], d[0]), dtype=np.float32)
        load().q2wref_get_embd(self._h, out.ctypes.data)
        return out

    def timings(self):
        t = (C.c_longlong * 4)()
        load().q2wref_timings(self._h, t)
        return dict(t_mel_us=t[0], t_encode_us=t[1], n_encode=t[2], t_load_us=t[3])

    def reset_timings(self):
        load().q2wref_reset_timings(self._h)


def ref_quantize(x: np.ndarray, ggml_type: int) -> np.ndarray:
    """ggml_quantize_chunk on float32 rows [nrows, K] -> raw bytes"""
    L = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    nrows, k = x.shape
    out = np.empty(L.q2wref_row_size(ggml_type, k) * nrows, dtype=np.uint8)
    L.q2wref_quantize(ggml_type, x.ctypes.data, out.ctypes.data, nrows, k)
    return out


def ref_dequantize(raw: np.ndarray, ggml_type: int, n: int) -> np.ndarray:
    L = load()
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(n, dtype=np.float32)
    L.q2wref_dequantize(ggml_type, raw.ctypes.data, out.ctypes.data, n)
    return out


def ref_gelu(x: np.ndarray) -> np.ndarray:
    L = load()
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    y = np.empty_like(x)
    L.q2wref_gelu(x.ctypes.data, y.ctypes.data, x.size)
    return y
