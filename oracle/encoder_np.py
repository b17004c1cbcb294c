"""oracle/encoder_np.py -- TEST INFRASTRUCTURE ONLY: numpy restatement of the reference encoder graph.

Follows /root/reference/src/qwen2-whisper.cpp:
  conv stem      whisper_build_graph_conv     :1892-1952  conv1d k3 s1 p1 + b, GELU, conv1d k3 s2 p1 + b, GELU
  encoder        whisper_build_graph_encoder  :1954-2203  + pos, 32 x {LN, QKV, softmax(QK^T)V, out-proj, +res, LN, fc1, GELU, fc2, +res},
                                                          avg-pool(2,2) over time, final LN
and the ggml CPU op semantics of SURVEY 2.3 (ggml/src/ggml.c): ggml_norm :11941-11990 (eps 1e-5), tanh-GELU
:2541-2547 evaluated through an F16 table :2556-2570, mul_mat with the activation converted to vec_dot_type
(F16 for F16 weights, Q8_0 blocks for Q8_0/Q4_0 weights :12507-12535), soft_max :13854-13940.

`mode`:
  "ggml"  mimic the CPU backend's roundings (activations -> F16 / Q8_0 before weight mat-muls, GELU via F16)   [default]
  "f32"   plain float32 math on the dequantised weights (the "dequantised-weight F32 restatement" of Appendix F)
Pinned against the reference itself by tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import numpy as np

GGML_TYPE_F32, GGML_TYPE_F16, GGML_TYPE_Q4_0, GGML_TYPE_Q8_0 = 0, 1, 2, 8


def _dequant(raw: np.ndarray, ttype: int, k: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw).view(np.uint8)
    if ttype == GGML_TYPE_F32:
        return raw.view(np.float32).reshape(-1, k).astype(np.float32)
    if ttype == GGML_TYPE_F16:
        return raw.view(np.float16).reshape(-1, k).astype(np.float32)
    if ttype == GGML_TYPE_Q8_0:     # dequantize_row_q8_0 ggml-quants.c:1616
        b = raw.reshape(-1, 34)
        d = b[:, :2].copy().view(np.float16).astype(np.float32)
        return (b[:, 2:].view(np.int8).astype(np.float32) * d).reshape(-1, k)
    if ttype == GGML_TYPE_Q4_0:     # dequantize_row_q4_0 ggml-quants.c:1522
        b = raw.reshape(-1, 18)
        d = b[:, :2].copy().view(np.float16).astype(np.float32)
        q = np.concatenate([(b[:, 2:] & 0xF).astype(np.int32) - 8, (b[:, 2:] >> 4).astype(np.int32) - 8], axis=1)
        return (q.astype(np.float32) * d).reshape(-1, k)
    raise ValueError(ttype)


def _q8_0_roundtrip(x: np.ndarray) -> np.ndarray:
    """quantize_row_q8_0 then dequantise: what the CPU backend feeds its int8 dot for Q8_0/Q4_0 weights.
    This is the vectorised x86 path the backend actually runs (ggml-quants.c:943-1000), which differs from the scalar
    _ref quantiser in two details: id = 127 / amax (not 1 / d) and round-to-nearest-even (not roundf)."""
    shp = x.shape
    xb = x.reshape(-1, 32).astype(np.float32)
    amax = np.abs(xb).max(axis=1)
    d = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(amax != 0, np.float32(127.0) / amax, np.float32(0)).astype(np.float32)
    q = np.rint((xb * inv[:, None]).astype(np.float32))
    d16 = d.astype(np.float16).astype(np.float32)
    return (q.astype(np.float32) * d16[:, None]).reshape(shp)


def gelu_tanh(x: np.ndarray, via_f16: bool) -> np.ndarray:
    """ggml_vec_gelu_f32 (ggml.c:2556-2570): x <= -10 -> 0, x >= 10 -> x, else table[f16(x)] with F16 table entries"""
    x = np.asarray(x, dtype=np.float32)
    xin = x.astype(np.float16).astype(np.float32) if via_f16 else x
    y = (0.5 * xin * (1.0 + np.tanh(np.float32(0.79788456080286535587989211986876) * xin * (1.0 + np.float32(0.044715) * xin * xin)))).astype(np.float32)
    if via_f16:
        y = y.astype(np.float16).astype(np.float32)
        y = np.where(x <= -10.0, np.float32(0.0), np.where(x >= 10.0, x, y)).astype(np.float32)
    return y


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=-1, keepdims=True)
    y = (x64 - mean)
    var = (y * y).mean(axis=-1, keepdims=True)
    scale = (1.0 / np.sqrt(var.astype(np.float32) + np.float32(eps))).astype(np.float32)
    return ((y.astype(np.float32) * scale) * g + b).astype(np.float32)


class EncoderOracle:
    def __init__(self, mf, mode: str = "ggml"):
        """mf: a parsed model file exposing .hparams, .tensor(name) -> (ttype, ne, data) records"""
        assert mode in ("ggml", "f32")
        self.mode = mode
        self.hp = mf.hparams
        self.w = {}
        self.t = {}
        for t in mf.tensors:
            self.w[t.name] = _dequant(t.data, t.ttype, t.ne[0]).reshape(tuple(reversed(t.ne)))
            self.t[t.name] = t.ttype

    def _matmul(self, x: np.ndarray, name: str) -> np.ndarray:
        w = self.w[name]
        if self.mode == "ggml":
            tt = self.t[name]
            if tt == GGML_TYPE_F16:
                x = x.astype(np.float16).astype(np.float32)
            elif tt in (GGML_TYPE_Q8_0, GGML_TYPE_Q4_0):
                x = _q8_0_roundtrip(x)
        return (x @ w.T).astype(np.float32)

    def conv_stem(self, mel_win: np.ndarray) -> np.ndarray:
        """mel_win [n_mel, 2*n_ctx] -> [n_ctx, D] (time-major, before the positional embedding)"""
        def conv(x, w, b, stride):   # x [C_in, L]; w [C_out, C_in, 3]; pad 1.  F32 x F32 (the harness upcasts F16 kernels)
            cin, L = x.shape
            xp = np.zeros((cin, L + 2), dtype=np.float32)
            xp[:, 1:-1] = x
            lo = (L + 2 - 3) // stride + 1
            cols = np.stack([xp[:, k:k + stride * lo:stride] for k in range(3)], axis=1)   # [C_in, 3, lo]
            y = w.reshape(w.shape[0], -1) @ cols.reshape(cin * 3, lo)
            return (y + b.reshape(-1, 1)).astype(np.float32)
        via = self.mode == "ggml"
        h = gelu_tanh(conv(mel_win.astype(np.float32), self.w["conv1.weight"], self.w["conv1.bias"], 1), via)
        h = gelu_tanh(conv(h, self.w["conv2.weight"], self.w["conv2.bias"], 2), via)
        return np.ascontiguousarray(h.T)

    def encode(self, mel_win: np.ndarray, return_pre_pool: bool = False, n_layers: int | None = None, taps: dict | None = None) -> np.ndarray:
        """n_layers: stop after that many encoder blocks (0 = conv stem + positional embedding); with return_pre_pool the residual
        stream at that point comes back -- `cur` / `inpL` of whisper_build_graph_encoder (:2005, :2154) -- for the stage tests.
        taps: a dict whose keys are layer counts; each is filled with a copy of the residual stream after that many blocks"""
        hp = self.hp
        T, D, H, L = hp["n_audio_ctx"], hp["n_audio_state"], hp["n_audio_head"], hp["n_audio_layer"]
        hd = D // H
        via = self.mode == "ggml"
        x = self.conv_stem(mel_win) + self.w["embed_positions.weight"][:T]
        scale = np.float32(1.0 / np.sqrt(float(hd)))
        if taps is not None and 0 in taps:
            taps[0] = x.copy()
        for i in range(L if n_layers is None else min(L, n_layers)):
            p = f"layers.{i}."
            c = layer_norm(x, self.w[p + "self_attn_layer_norm.weight"], self.w[p + "self_attn_layer_norm.bias"])
            q = (self._matmul(c, p + "self_attn.q_proj.weight") + self.w[p + "self_attn.q_proj.bias"]) * scale
            k = self._matmul(c, p + "self_attn.k_proj.weight")
            v = self._matmul(c, p + "self_attn.v_proj.weight") + self.w[p + "self_attn.v_proj.bias"]
            qh = q.reshape(T, H, hd).transpose(1, 0, 2)
            kh = k.reshape(T, H, hd).transpose(1, 0, 2)
            vh = v.reshape(T, H, hd).transpose(1, 0, 2)
            s = qh @ kh.transpose(0, 2, 1)                                  # F32 x F32, no casts (:2082-2100)
            s = s - s.max(axis=-1, keepdims=True)
            e = np.exp(s).astype(np.float32)
            pr = (e / e.astype(np.float64).sum(axis=-1, keepdims=True)).astype(np.float32)
            a = (pr @ vh).transpose(1, 0, 2).reshape(T, D)
            x = x + self._matmul(a, p + "self_attn.out_proj.weight") + self.w[p + "self_attn.out_proj.bias"]
            c = layer_norm(x, self.w[p + "final_layer_norm.weight"], self.w[p + "final_layer_norm.bias"])
            hmid = gelu_tanh(self._matmul(c, p + "fc1.weight") + self.w[p + "fc1.bias"], via)
            x = x + self._matmul(hmid, p + "fc2.weight") + self.w[p + "fc2.bias"]
            if taps is not None and (i + 1) in taps:
                taps[i + 1] = x.copy()
        if return_pre_pool:
            return x
        y = ((x[0::2] + x[1::2]) / np.float32(2.0)).astype(np.float32)      # ggml_pool_1d AVG k2 s2 (:2165)
        return layer_norm(y, self.w["layer_norm.weight"], self.w["layer_norm.bias"])
