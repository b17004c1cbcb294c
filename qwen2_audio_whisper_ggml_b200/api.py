"""Python mirror of the reference's public API (include/qwen2-whisper.h) over libq2w_b200.so.

Same names and argument meaning as the C functions so the parity tests read like calls against the reference:
``Context.init_from_file`` / ``init_from_buffer`` -> whisper_init_from_*_with_params, ``pcm_to_mel``, ``set_mel``,
``encode``, ``full``, ``n_len``, ``print_timings`` ... plus the additive accessors (``get_mel``, ``get_embeddings``,
``encode_batch``).  Nothing here computes: every method is one C call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l


class ContextParams(C.Structure):        # struct whisper_context_params, include/qwen2-whisper.h
    class _Aheads(C.Structure):
        _fields_ = [("n_heads", C.c_size_t), ("heads", C.c_void_p)]
    _fields_ = [("use_gpu", C.c_bool), ("flash_attn", C.c_bool), ("gpu_device", C.c_int),
                ("dtw_token_timestamps", C.c_bool), ("dtw_aheads_preset", C.c_int), ("dtw_n_top", C.c_int),
                ("dtw_aheads", _Aheads), ("dtw_mem_size", C.c_size_t)]


class FullParams(C.Structure):           # struct whisper_full_params, include/qwen2-whisper.h
    _fields_ = [("n_threads", C.c_int), ("n_max_text_ctx", C.c_int), ("offset_ms", C.c_int), ("duration_ms", C.c_int),
                ("translate", C.c_bool), ("no_context", C.c_bool), ("no_timestamps", C.c_bool), ("single_segment", C.c_bool),
                ("print_special", C.c_bool), ("print_progress", C.c_bool), ("print_realtime", C.c_bool), ("print_timestamps", C.c_bool),
                ("token_timestamps", C.c_bool), ("thold_pt", C.c_float), ("thold_ptsum", C.c_float), ("max_len", C.c_int),
                ("split_on_word", C.c_bool), ("max_tokens", C.c_int), ("debug_mode", C.c_bool), ("audio_ctx", C.c_int),
                ("tdrz_enable", C.c_bool), ("suppress_regex", C.c_char_p), ("initial_prompt", C.c_char_p),
                ("prompt_tokens", C.c_void_p), ("prompt_n_tokens", C.c_int), ("language", C.c_char_p), ("detect_language", C.c_bool),
                ("suppress_blank", C.c_bool), ("suppress_non_speech_tokens", C.c_bool),
                ("temperature", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
                ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float), ("no_speech_thold", C.c_float),
                ("new_segment_callback", C.c_void_p), ("new_segment_callback_user_data", C.c_void_p),
                ("progress_callback", C.c_void_p), ("progress_callback_user_data", C.c_void_p),
                ("encoder_begin_callback", C.c_void_p), ("encoder_begin_callback_user_data", C.c_void_p),
                ("abort_callback", C.c_void_p), ("abort_callback_user_data", C.c_void_p), ("i_start_rule", C.c_size_t)]


LOG_CB = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_WSIGS = {
    "whisper_context_default_params": (ContextParams, []),
    "whisper_full_default_params": (FullParams, []),
    "whisper_init_from_file_with_params": (_vp, [C.c_char_p, ContextParams]),
    "whisper_init_from_buffer_with_params": (_vp, [_vp, _sz, ContextParams]),
    "whisper_init_from_file_with_params_no_state": (_vp, [C.c_char_p, ContextParams]),
    "whisper_init_from_buffer_with_params_no_state": (_vp, [_vp, _sz, ContextParams]),
    "whisper_init_state": (_vp, [_vp]),
    "whisper_free": (None, [_vp]),
    "whisper_free_state": (None, [_vp]),
    "whisper_pcm_to_mel": (_i, [_vp, _vp, _i, _i]),
    "whisper_pcm_to_mel_with_state": (_i, [_vp, _vp, _vp, _i, _i]),
    "whisper_set_mel": (_i, [_vp, _vp, _i, _i]),
    "whisper_encode": (_i, [_vp, _i, _i]),
    "whisper_encode_with_state": (_i, [_vp, _vp, _i, _i]),
    "whisper_full": (_i, [_vp, FullParams, _vp, _i]),
    "whisper_full_with_state": (_i, [_vp, _vp, FullParams, _vp, _i]),
    "whisper_n_len": (_i, [_vp]),
    "whisper_n_len_from_state": (_i, [_vp]),
    "whisper_model_n_vocab": (_i, [_vp]), "whisper_model_n_audio_ctx": (_i, [_vp]), "whisper_model_n_audio_state": (_i, [_vp]),
    "whisper_model_n_audio_head": (_i, [_vp]), "whisper_model_n_audio_layer": (_i, [_vp]), "whisper_model_n_mels": (_i, [_vp]),
    "whisper_model_ftype": (_i, [_vp]), "whisper_model_type": (_i, [_vp]),
    "whisper_print_timings": (None, [_vp]), "whisper_reset_timings": (None, [_vp]),
    "whisper_print_system_info": (C.c_char_p, []),
    "whisper_log_set": (None, [LOG_CB, _vp]),
    "whisper_print_emb_enc": (None, [_vp]),
    "whisper_embd_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "whisper_get_embeddings": (_i, [_vp, _vp, _sz]),
    "whisper_get_embeddings_from_state": (_i, [_vp, _vp, _sz]),
    "whisper_get_embeddings_device": (_vp, [_vp]),
    "whisper_get_mel": (_i, [_vp, _vp, _sz]),
    "whisper_get_mel_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "whisper_encode_batch": (_i, [_vp, _vp, _sz, _vp, _i, _vp]),
    "whisper_encode_batch_device": (_i, [_vp, _vp, _sz, _vp, _i]),
    "whisper_set_max_batch": (_i, [_vp, _i]),
    "whisper_encode_offsets": (_i, [_vp, _vp, _i, _vp]),
    "whisper_encode_batch_async": (_i, [_vp, _vp, C.c_size_t, _vp, _i, _vp]),
    "whisper_encode_batch_wait": (_i, [_vp, _i]),
    "whisper_q2w_state": (_vp, [_vp]),
    "whisper_init_from_file": (_vp, [C.c_char_p]),
    "whisper_init_from_buffer": (_vp, [_vp, _sz]),
    "whisper_init_from_file_no_state": (_vp, [C.c_char_p]),
    "whisper_init_from_buffer_no_state": (_vp, [_vp, _sz]),
    "whisper_init_from_file_multi": (_vp, [C.c_char_p, ContextParams, C.POINTER(_i), _i]),
    "whisper_init_from_buffer_multi": (_vp, [_vp, _sz, ContextParams, C.POINTER(_i), _i]),
    "whisper_n_devices": (_i, [_vp]),
    "whisper_device": (_i, [_vp, _i]),
    "whisper_encode_batch_multi": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _i]),
    "whisper_get_gathered_device": (_vp, [_vp]),
    "whisper_q2w_multi": (_vp, [_vp]),
    "whisper_set_projector": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "whisper_project": (_i, [_vp, _vp, _sz]),
    "whisper_projection_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
}
_bound = False


def wlib() -> C.CDLL:
    global _bound
    lib = _l.load_library()
    if not _bound:
        for name, (res, args) in _WSIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return lib


def default_context_params() -> ContextParams:
    return wlib().whisper_context_default_params()


_log_keepalive = []


def log_set(fn=None):
    """whisper_log_set: fn(level:int, text:str) or None to restore stderr logging."""
    if fn is None:
        wlib().whisper_log_set(C.cast(None, LOG_CB), None)
        return
    cb = LOG_CB(lambda lvl, txt, ud: fn(lvl, txt.decode("utf-8", "replace")))
    _log_keepalive.append(cb)
    wlib().whisper_log_set(cb, None)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Context:
    """struct whisper_context * with the reference's lifetime rules (whisper_free releases the default state)."""

    def __init__(self, handle: int):
        if not handle:
            raise _l.Q2WError(-1, "whisper_init_* returned NULL (see log)")
        self._h = handle

    # ---- init / free
    @classmethod
    def init_from_file(cls, path: str, params: ContextParams | None = None) -> "Context":
        p = params if params is not None else default_context_params()
        return cls(wlib().whisper_init_from_file_with_params(path.encode(), p))

    @classmethod
    def init_from_buffer(cls, buf: bytes, params: ContextParams | None = None, devices=None) -> "Context":
        """devices: explicit list of CUDA ordinals, one weight replica each (whisper_init_from_buffer_multi); None -> params.gpu_device
        (-1 = every visible sm_100 device)"""
        p = params if params is not None else default_context_params()
        raw = (C.c_char * len(buf)).from_buffer_copy(buf)
        if devices is not None:
            d = (_i * len(devices))(*devices)
            return cls(wlib().whisper_init_from_buffer_multi(C.cast(raw, C.c_void_p), len(buf), p, d, len(devices)))
        return cls(wlib().whisper_init_from_buffer_with_params(C.cast(raw, C.c_void_p), len(buf), p))

    def free(self):
        if self._h:
            wlib().whisper_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- reference API
    def pcm_to_mel(self, samples, n_threads: int = 1) -> int:
        s = _f32(samples)
        return wlib().whisper_pcm_to_mel(self._h, s.ctypes.data, s.size, n_threads)

    def set_mel(self, data, n_len: int, n_mel: int) -> int:
        d = _f32(data)
        return wlib().whisper_set_mel(self._h, d.ctypes.data, n_len, n_mel)

    def encode(self, offset: int = 0, n_threads: int = 1) -> int:
        return wlib().whisper_encode(self._h, offset, n_threads)

    def full(self, samples=None, params: FullParams | None = None, offset_ms: int = 0) -> int:
        p = params if params is not None else wlib().whisper_full_default_params()
        if offset_ms:
            p.offset_ms = offset_ms
        if samples is None:
            return wlib().whisper_full(self._h, p, None, 0)
        s = _f32(samples)
        return wlib().whisper_full(self._h, p, s.ctypes.data, s.size)

    def n_len(self) -> int:
        return wlib().whisper_n_len(self._h)

    def model_n(self, what: str) -> int:
        return getattr(wlib(), f"whisper_model_{what}")(self._h)

    def print_timings(self):
        wlib().whisper_print_timings(self._h)

    def reset_timings(self):
        wlib().whisper_reset_timings(self._h)

    def print_emb_enc(self):
        wlib().whisper_print_emb_enc(self._h)

    # ---- additive API
    def embd_dims(self):
        a, b, c = _i(), _i(), _i()
        if wlib().whisper_embd_dims(self._h, C.byref(a), C.byref(b), C.byref(c)) != 0:
            raise _l.Q2WError(-1, "whisper_embd_dims failed")
        return a.value, b.value, c.value

    def get_embeddings(self) -> np.ndarray:
        nw, no, ns = self.embd_dims()
        out = np.empty((nw, no, ns), dtype=np.float32)
        if wlib().whisper_get_embeddings(self._h, out.ctypes.data, out.size) != 0:
            raise _l.Q2WError(-1, "whisper_get_embeddings failed (see log)")
        return out

    def mel_dims(self):
        """(n_len, n_len_org, n_mel) of the default state's mel"""
        a, b, c = _i(), _i(), _i()
        if wlib().whisper_get_mel_dims(self._h, C.byref(a), C.byref(b), C.byref(c)) != 0:
            raise _l.Q2WError(-1, "whisper_get_mel_dims failed")
        return a.value, b.value, c.value

    def get_mel(self) -> np.ndarray:
        n_len, _, n_mel = self.mel_dims()
        out = np.empty((n_mel, n_len), dtype=np.float32)
        if wlib().whisper_get_mel(self._h, out.ctypes.data, out.size) != 0:
            raise _l.Q2WError(-1, "whisper_get_mel failed (see log)")
        return out

    def set_max_batch(self, n: int) -> int:
        return wlib().whisper_set_max_batch(self._h, n)

    def encode_batch(self, windows, n_samples=None, out: np.ndarray | None = None, want_host: bool = True):
        """windows: float32 [B, stride] host array (or a pinned torch tensor's numpy view)."""
        w = windows if (isinstance(windows, np.ndarray) and windows.dtype == np.float32 and windows.flags.c_contiguous) else _f32(windows)
        B, stride = w.shape
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, dtype=np.int32)
        if want_host and out is None:
            out = np.empty((B, self.model_n("n_audio_ctx") // 2, self.model_n("n_audio_state")), dtype=np.float32)
        rc = wlib().whisper_encode_batch(self._h, w.ctypes.data, stride, None if ns is None else ns.ctypes.data, B,
                                         out.ctypes.data if out is not None else None)
        if rc != 0:
            raise _l.Q2WError(rc, "whisper_encode_batch failed (see log)")
        return out

    def n_devices(self) -> int:
        return wlib().whisper_n_devices(self._h)

    def devices(self) -> list[int]:
        return [wlib().whisper_device(self._h, i) for i in range(self.n_devices())]

    def encode_batch_multi(self, windows, n_samples=None, out: np.ndarray | None = None, gather_device: int = -1, want_host: bool = True):
        """whisper_encode_batch_multi: shard over every replica of the context; gather_device >= 0 also assembles all embeddings on that
        device (gathered_device_ptr / gathered())"""
        w = windows if (isinstance(windows, np.ndarray) and windows.dtype == np.float32 and windows.flags.c_contiguous) else _f32(windows)
        B, stride = w.shape
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, dtype=np.int32)
        if want_host and out is None:
            out = np.empty((B, self.model_n("n_audio_ctx") // 2, self.model_n("n_audio_state")), dtype=np.float32)
        rc = wlib().whisper_encode_batch_multi(self._h, w.ctypes.data, stride, None if ns is None else ns.ctypes.data, B,
                                               out.ctypes.data if out is not None else None, gather_device)
        if rc != 0:
            raise _l.Q2WError(rc, "whisper_encode_batch_multi failed (see log)")
        return out

    def gathered_device_ptr(self) -> int:
        return wlib().whisper_get_gathered_device(self._h)

    def gathered(self, B: int) -> np.ndarray:
        """host copy of the embeddings gathered on one device by the last encode_batch_multi(..., gather_device=k)"""
        mm = wlib().whisper_q2w_multi(self._h)
        out = np.empty((B, self.model_n("n_audio_ctx") // 2, self.model_n("n_audio_state")), dtype=np.float32)
        if not mm:
            raise _l.Q2WError(-1, "single-device context: nothing was gathered")
        _l.check(_l.load_library().q2w_multi_get_gathered(mm, out.ctypes.data, out.size))
        return out

    def encode_batch_async(self, windows: np.ndarray, out: np.ndarray, n_samples=None) -> int:
        """Queue a batch and return a ticket; `windows` and `out` (float32, C-contiguous, ideally pinned) must stay alive and untouched
        until wait(ticket). Two batches may be in flight: submit i + 1, then wait for i."""
        assert windows.dtype == np.float32 and windows.flags.c_contiguous and out.dtype == np.float32 and out.flags.c_contiguous
        B, stride = windows.shape
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, dtype=np.int32)
        t = wlib().whisper_encode_batch_async(self._h, windows.ctypes.data, stride, None if ns is None else ns.ctypes.data, B, out.ctypes.data)
        if t < 0:
            raise _l.Q2WError(t, "whisper_encode_batch_async failed (see log)")
        return t

    def wait(self, ticket: int) -> None:
        if wlib().whisper_encode_batch_wait(self._h, ticket) != 0:
            raise _l.Q2WError(-1, "whisper_encode_batch_wait failed (see log)")

    def encode_long(self, samples, out: np.ndarray | None = None) -> np.ndarray:
        """Audio of any length: cut into 30 s windows (the last one ragged), each with its own mel normalisation, and run
        them as one batch (BASELINE config 5; the fork itself never iterates past the first window)."""
        s = _f32(samples).reshape(-1)
        win = 2 * self.model_n("n_audio_ctx") * 160
        nw = max(1, -(-s.size // win))
        padded = np.zeros((nw, win), dtype=np.float32)
        padded.reshape(-1)[: s.size] = s
        ns = np.full(nw, win, dtype=np.int32)
        ns[-1] = s.size - (nw - 1) * win
        return self.encode_batch(padded, ns, out=out)

    def encode_stream(self, samples, hop_frames: int | None = None) -> np.ndarray:
        """Whole-file streaming: ONE mel over the full audio (global normalisation, like the reference's whisper_pcm_to_mel),
        then every window starting at k * hop_frames (default: 2 * n_audio_ctx, i.e. back-to-back 30 s windows) as one batch."""
        if self.pcm_to_mel(samples) != 0:
            raise _l.Q2WError(-1, "whisper_pcm_to_mel failed")
        win = 2 * self.model_n("n_audio_ctx")
        hop = hop_frames or win
        n_org = self.n_len()
        offs = np.arange(0, max(n_org, 1), hop, dtype=np.int32)
        out = np.empty((offs.size, self.model_n("n_audio_ctx") // 2, self.model_n("n_audio_state")), dtype=np.float32)
        if wlib().whisper_encode_offsets(self._h, offs.ctypes.data, offs.size, out.ctypes.data) != 0:
            raise _l.Q2WError(-1, "whisper_encode_offsets failed (see log)")
        return out

    def encode_batch_device(self, dev_ptr: int, stride: int, B: int, n_samples=None) -> int:
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, dtype=np.int32)
        return wlib().whisper_encode_batch_device(self._h, dev_ptr, stride, None if ns is None else ns.ctypes.data, B)

    def q2w_state(self) -> int:
        return wlib().whisper_q2w_state(self._h)

    def set_projector(self, weight: np.ndarray, bias: np.ndarray | None = None) -> int:
        """multi_modal_projector.linear: weight [n_out, n_audio_state] float32 or float16, bias [n_out] float32"""
        w = np.ascontiguousarray(weight)
        assert w.dtype in (np.float32, np.float16) and w.ndim == 2
        b = None if bias is None else _f32(bias)
        return wlib().whisper_set_projector(self._h, 0 if w.dtype == np.float32 else 1, w.shape[0], w.ctypes.data, w.nbytes,
                                            None if b is None else b.ctypes.data)

    def project(self) -> np.ndarray:
        """projector applied to the embeddings of the last encode: [n_windows, n_audio_ctx // 2, n_out]"""
        nw, no, _ = self.embd_dims()
        a, b = _i(), _i()
        if wlib().whisper_projection_dims(self._h, C.byref(a), C.byref(b)) != 0 or b.value <= 0:
            raise _l.Q2WError(-1, "no projector set")
        out = np.empty((nw, no, b.value), dtype=np.float32)
        if wlib().whisper_project(self._h, out.ctypes.data, out.size) != 0:
            raise _l.Q2WError(-1, "whisper_project failed: " + _l.load_library().q2w_last_error().decode("utf-8", "replace"))
        return out

    def debug_forward_layers(self, n_layers: int) -> None:
        """stage tap (parity tests): every following forward stops after n_layers encoder blocks (-1 = all)"""
        _l.check(_l.load_library().q2w_debug_forward_layers(self.q2w_state(), n_layers))

    def debug_residual(self, window: int = 0) -> np.ndarray:
        """F32 residual stream [n_audio_ctx, n_audio_state] of one window of the last forward"""
        out = np.empty((self.model_n("n_audio_ctx"), self.model_n("n_audio_state")), dtype=np.float32)
        _l.check(_l.load_library().q2w_debug_get_residual(self.q2w_state(), window, out.ctypes.data))
        return out

    def embeddings_device_ptr(self) -> int:
        return wlib().whisper_get_embeddings_device(self._h)
