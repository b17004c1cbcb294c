"""ggml block quantisation used by the model-file tooling (the fork's `quantize` path).

Restates, for the three on-path types only, what examples/common-ggml.cpp:41 (ggml_common_quantize_0) does through
ggml_quantize_chunk: row-wise blocks of 32 along ne[0].
  block_q8_0 {f16 d; int8 qs[32]}  34 B   ggml/src/ggml-common.h:186-191   quantize_row_q8_0_ref ggml-quants.c:848-871
  block_q4_0 {f16 d; u8  qs[16]}   18 B   ggml/src/ggml-common.h:144-148   quantize_row_q4_0_ref ggml-quants.c:668-703
tests/test_host_cpu.py pins these bit-for-bit against the reference's own ggml_quantize_chunk: live (oracle/_ref,
test_quantisers_bit_exact_vs_ggml) and against blocks it produced once (tests/golden/ggml_quant_blocks.npz).
"""
from __future__ import annotations

import numpy as np

GGML_TYPE_F32, GGML_TYPE_F16, GGML_TYPE_Q4_0, GGML_TYPE_Q8_0 = 0, 1, 2, 8
GGML_FTYPE = {GGML_TYPE_F32: 0, GGML_TYPE_F16: 1, GGML_TYPE_Q4_0: 2, GGML_TYPE_Q8_0: 7}   # enum ggml_ftype, ggml.h
GGML_QNT_VERSION = 2
QK = 32
BLOCK_BYTES = {GGML_TYPE_Q8_0: 34, GGML_TYPE_Q4_0: 18}
TYPE_NAMES = {GGML_TYPE_F32: "f32", GGML_TYPE_F16: "f16", GGML_TYPE_Q4_0: "q4_0", GGML_TYPE_Q8_0: "q8_0"}


def row_bytes(ggml_type: int, ne0: int) -> int:
    if ggml_type == GGML_TYPE_F32:
        return 4 * ne0
    if ggml_type == GGML_TYPE_F16:
        return 2 * ne0
    assert ne0 % QK == 0
    return ne0 // QK * BLOCK_BYTES[ggml_type]


def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    """float32 [..., K] -> uint8 [..., K/32*34]"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    k = x.shape[-1]
    assert k % QK == 0
    xb = x.reshape(-1, QK)
    amax = np.max(np.abs(xb), axis=1)
    d = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    x0 = (xb * inv[:, None]).astype(np.float32).astype(np.float64)
    q = np.trunc(x0 + np.copysign(0.5, x0)).astype(np.int8)            # roundf: half away from zero
    out = np.empty((xb.shape[0], 34), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out.reshape(*x.shape[:-1], k // QK * 34)


def dequantize_q8_0(raw: np.ndarray, k: int) -> np.ndarray:
    b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, 34)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)           # [nb, 1]
    q = b[:, 2:].view(np.int8).astype(np.float32)
    return (q * d).reshape(-1, k)


def quantize_q4_0(x: np.ndarray) -> np.ndarray:
    """float32 [..., K] -> uint8 [..., K/32*18]"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    k = x.shape[-1]
    assert k % QK == 0
    xb = x.reshape(-1, QK)
    idx = np.argmax(np.abs(xb), axis=1)                                # first occurrence of the max magnitude
    mx = xb[np.arange(xb.shape[0]), idx]
    d = (mx / np.float32(-8.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    xs = (xb * inv[:, None]).astype(np.float32)
    v = (xs + np.float32(8.5)).astype(np.float32)
    qi = np.minimum(15, np.trunc(v).astype(np.int64).astype(np.int8)).astype(np.uint8)   # MIN(15, (int8_t)(x + 8.5f))
    out = np.empty((xb.shape[0], 18), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = qi[:, :16] | (qi[:, 16:] << 4)
    return out.reshape(*x.shape[:-1], k // QK * 18)


def dequantize_q4_0(raw: np.ndarray, k: int) -> np.ndarray:
    b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, 18)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)
    qs = b[:, 2:]
    lo = (qs & 0x0F).astype(np.int32) - 8
    hi = (qs >> 4).astype(np.int32) - 8
    q = np.concatenate([lo, hi], axis=1).astype(np.float32)
    return (q * d).reshape(-1, k)


def quantize(x: np.ndarray, ggml_type: int) -> np.ndarray:
    """float32 rows -> raw ggml bytes (uint8) for F32 / F16 / Q8_0 / Q4_0"""
    if ggml_type == GGML_TYPE_F32:
        return np.ascontiguousarray(x, dtype=np.float32).view(np.uint8)
    if ggml_type == GGML_TYPE_F16:
        return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).view(np.uint8)
    if ggml_type == GGML_TYPE_Q8_0:
        return quantize_q8_0(x)
    if ggml_type == GGML_TYPE_Q4_0:
        return quantize_q4_0(x)
    raise ValueError(f"ggml type {ggml_type} is not on this path")


def dequantize(raw: np.ndarray, ggml_type: int, k: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw).view(np.uint8)
    if ggml_type == GGML_TYPE_F32:
        return raw.view(np.float32).reshape(-1, k).copy()
    if ggml_type == GGML_TYPE_F16:
        return raw.view(np.float16).astype(np.float32).reshape(-1, k)
    if ggml_type == GGML_TYPE_Q8_0:
        return dequantize_q8_0(raw, k)
    if ggml_type == GGML_TYPE_Q4_0:
        return dequantize_q4_0(raw, k)
    raise ValueError(f"ggml type {ggml_type} is not on this path")
