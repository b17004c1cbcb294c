"""ctypes binding of libq2w_b200.so -- declares every symbol of include/q2w_b200.h."""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libq2w_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "q2w_b200.h")

Q2W_TYPE_F32, Q2W_TYPE_F16, Q2W_TYPE_Q4_0, Q2W_TYPE_Q8_0 = 0, 1, 2, 8
EPI_BIAS_F16, EPI_BIAS_GELU_F16, EPI_BIAS_RESID_F32, EPI_BIAS_GELU_POS_F32, EPI_BIAS_F32 = range(5)


class Q2WError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"q2w error {code}: {msg}")
        self.code = code


class HParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
        "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype")]


_lock = threading.Lock()
_lib = None

_vp, _i, _sz, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_SIGS = {
    "q2w_model_create": (_i, [C.POINTER(_vp), C.POINTER(HParams), _i, _i]),
    "q2w_model_upload_filters": (_i, [_vp, _vp, _i, _i]),
    "q2w_model_upload_tensor": (_i, [_vp, C.c_char_p, _i, _i, C.POINTER(C.c_int32), _vp, _sz]),
    "q2w_model_tensor_bytes": (_sz, [_vp, C.c_char_p]),
    "q2w_model_finalize": (_i, [_vp]),
    "q2w_model_free": (None, [_vp]),
    "q2w_model_n_tensors_expected": (_i, [_vp]),
    "q2w_model_n_tensors_loaded": (_i, [_vp]),
    "q2w_model_weight_bytes": (_sz, [_vp]),
    "q2w_state_create": (_i, [C.POINTER(_vp), _vp, _i]),
    "q2w_state_free": (None, [_vp]),
    "q2w_pcm_to_mel": (_i, [_vp, _vp, _i]),
    "q2w_set_mel": (_i, [_vp, _vp, _i, _i]),
    "q2w_mel_n_len": (_i, [_vp]),
    "q2w_mel_n_len_org": (_i, [_vp]),
    "q2w_get_mel": (_i, [_vp, _vp, _sz]),
    "q2w_encode": (_i, [_vp, _i]),
    "q2w_encode_batch_host": (_i, [_vp, _vp, _sz, _vp, _i, _vp]),
    "q2w_encode_batch_device": (_i, [_vp, _vp, _sz, _vp, _i]),
    "q2w_encode_offsets": (_i, [_vp, _vp, _i, _vp]),
    "q2w_encode_batch_host_async": (_i, [_vp, _vp, C.c_size_t, _vp, _i, _vp, _vp]),
    "q2w_encode_batch_wait": (_i, [_vp, _i]),
    "q2w_embd_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "q2w_get_embeddings": (_i, [_vp, _vp, _sz, _sz]),
    "q2w_embeddings_device": (_vp, [_vp]),
    "q2w_get_batch_mel": (_i, [_vp, _i, _vp]),
    "q2w_get_timings": (None, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "q2w_reset_timings": (None, [_vp]),
    "q2w_profile_enable": (_i, [_vp, _i]),
    "q2w_profile_read": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "q2w_state_stream": (_vp, [_vp]),
    "q2w_sync": (_i, [_vp]),
    "q2w_last_error": (C.c_char_p, []),
    "q2w_kernel_launches": (C.c_long, []),
    "q2w_device_count": (_i, []),
    "q2w_build_info": (C.c_char_p, []),
    "q2w_op_gemm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _f, _vp]),
    "q2w_op_gemm_q": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _f, _vp]),
    "q2w_op_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "q2w_op_pool_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "q2w_op_attention": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "q2w_op_conv1_operand": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "q2w_op_set_gemm_splitk": (None, [_i]),
    "q2w_debug_forward_layers": (_i, [_vp, _i]),
    "q2w_debug_get_residual": (_i, [_vp, _i, _vp]),
    "q2w_model_device": (_i, [_vp]),
    "q2w_model_upload_projector": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "q2w_model_projector_width": (_i, [_vp]),
    "q2w_project": (_i, [_vp, _vp, _sz]),
    "q2w_projection_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "q2w_projected_device": (_vp, [_vp]),
    "q2w_multi_create": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _i]),
    "q2w_multi_free": (None, [_vp]),
    "q2w_multi_n_devices": (_i, [_vp]),
    "q2w_multi_device": (_i, [_vp, _i]),
    "q2w_multi_state": (_vp, [_vp, _i]),
    "q2w_multi_shard_bounds": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "q2w_multi_set_max_batch": (_i, [_vp, _i]),
    "q2w_multi_encode_batch_host": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _i]),
    "q2w_multi_encode_batch_host_async": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp]),
    "q2w_multi_encode_batch_wait": (_i, [_vp, _i]),
    "q2w_multi_gathered_device": (_vp, [_vp]),
    "q2w_multi_get_gathered": (_i, [_vp, _vp, _sz]),
    "q2w_multi_last_device_ms": (C.c_double, [_vp, _i]),
    "q2w_state_set_max_batch": (_i, [_vp, _i]),
    "q2w_state_max_batch": (_i, [_vp]),
    "q2w_op_dequant": (_i, [_vp, _i, _vp, _sz, _i, _vp]),
    "q2w_op_dequant_multi": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "q2w_op_conv2_im2col": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "q2w_op_mel": (_i, [_vp, _i, _vp, _sz, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
}


def header_symbols(path: str = HEADER_PATH) -> list[str]:
    """Every Q2W_API function name the public header declares."""
    src = open(path).read()
    return sorted(set(re.findall(r"Q2W_API\s+[\w\s\*]+?\b(q2w_\w+)\s*\(", src)))


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built (there is no fallback path)."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise Q2WError(-3, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(make -C qwen2_audio_whisper_ggml_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(p)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        raise Q2WError(rc, load_library().q2w_last_error().decode("utf-8", "replace"))
