"""Model-file tooling for the fork's ggml `.bin` format.

Reader + writer of the exact byte stream src/qwen2-whisper.cpp:1350-1872 parses and models/convert-pt-to-ggml.py:266-339
emits: magic, 11 int32 hparams, mel filterbank, vocab, then {n_dims, name_len, ttype, ne[], name, data} records.
`quantize_model` is the equivalent of the fork's quantiser (examples/common-ggml.cpp:41): 2-D tensors become
Q8_0 / Q4_0 blocks, except embed_positions.weight and the (2-D!) conv biases; 3-D conv kernels stay F16, 1-D stay F32.
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass, field

import numpy as np

from . import ggml_quant as gq

GGML_FILE_MAGIC = 0x67676D6C
HPARAM_NAMES = ("n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
                "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype")
NO_QUANT = ("embed_positions.weight", "conv1.bias", "conv2.bias")


@dataclass
class TensorRec:
    name: str
    ttype: int
    ne: tuple          # ggml order: innermost first
    data: np.ndarray   # raw bytes, uint8

    @property
    def nbytes(self) -> int:
        return int(self.data.size)


@dataclass
class ModelFile:
    hparams: dict
    filters: np.ndarray                      # float32 [n_mel, n_fft]
    vocab: list = field(default_factory=list)  # list[bytes]
    tensors: list = field(default_factory=list)  # list[TensorRec]

    @property
    def wtype(self) -> int:
        ft = self.hparams["ftype"] % 1000
        return {0: gq.GGML_TYPE_F32, 1: gq.GGML_TYPE_F16, 2: gq.GGML_TYPE_Q4_0, 7: gq.GGML_TYPE_Q8_0}[ft]

    def tensor(self, name: str) -> TensorRec:
        for t in self.tensors:
            if t.name == name:
                return t
        raise KeyError(name)

    def as_float(self, name: str) -> np.ndarray:
        """tensor as float32 in numpy (outermost-first) shape"""
        t = self.tensor(name)
        rows = gq.dequantize(t.data, t.ttype, t.ne[0])
        return rows.reshape(tuple(reversed(t.ne)))


def tensor_names(n_layer: int) -> list[str]:
    names = ["embed_positions.weight", "conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
             "layer_norm.weight", "layer_norm.bias"]
    for i in range(n_layer):
        p = f"layers.{i}."
        names += [p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias",
                  p + "self_attn.q_proj.weight", p + "self_attn.q_proj.bias", p + "self_attn.k_proj.weight",
                  p + "self_attn.v_proj.weight", p + "self_attn.v_proj.bias",
                  p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                  p + "final_layer_norm.weight", p + "final_layer_norm.bias",
                  p + "fc1.weight", p + "fc1.bias", p + "fc2.weight", p + "fc2.bias"]
    return names


def expected_shapes(hp: dict) -> dict:
    """name -> ggml ne tuple (src/qwen2-whisper.cpp:1591-1638)"""
    D, T, NM, L = hp["n_audio_state"], hp["n_audio_ctx"], hp["n_mels"], hp["n_audio_layer"]
    sh = {"embed_positions.weight": (D, T), "conv1.weight": (3, NM, D), "conv1.bias": (1, D),
          "conv2.weight": (3, D, D), "conv2.bias": (1, D), "layer_norm.weight": (D,), "layer_norm.bias": (D,)}
    for i in range(L):
        p = f"layers.{i}."
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sh[p + n + ".weight"] = (D,)
            sh[p + n + ".bias"] = (D,)
        for n in ("q_proj", "v_proj", "out_proj"):
            sh[p + f"self_attn.{n}.weight"] = (D, D)
            sh[p + f"self_attn.{n}.bias"] = (D,)
        sh[p + "self_attn.k_proj.weight"] = (D, D)
        sh[p + "fc1.weight"] = (D, 4 * D)
        sh[p + "fc1.bias"] = (4 * D,)
        sh[p + "fc2.weight"] = (4 * D, D)
        sh[p + "fc2.bias"] = (D,)
    return sh


def tensor_type_for(name: str, ne: tuple, wtype: int) -> int:
    """ggml type the loader expects for `name` in a file of weight type `wtype` (:1542-1543, :1591-1638)"""
    if len(ne) == 3:                      # conv kernels: vtype
        return gq.GGML_TYPE_F32 if wtype == gq.GGML_TYPE_F32 else gq.GGML_TYPE_F16
    if len(ne) == 2 and name not in NO_QUANT:
        return wtype
    return gq.GGML_TYPE_F32


def write_model(f, mf: ModelFile) -> None:
    f.write(struct.pack("<i", GGML_FILE_MAGIC))
    f.write(struct.pack("<11i", *[mf.hparams[k] for k in HPARAM_NAMES]))
    filt = np.ascontiguousarray(mf.filters, dtype=np.float32)
    f.write(struct.pack("<ii", filt.shape[0], filt.shape[1]))
    f.write(filt.tobytes())
    f.write(struct.pack("<i", len(mf.vocab)))
    for w in mf.vocab:
        f.write(struct.pack("<I", len(w)))
        f.write(w)
    for t in mf.tensors:
        nb = t.name.encode()
        f.write(struct.pack("<iii", len(t.ne), len(nb), t.ttype))
        f.write(struct.pack(f"<{len(t.ne)}i", *t.ne))
        f.write(nb)
        f.write(np.ascontiguousarray(t.data).view(np.uint8).tobytes())


def to_bytes(mf: ModelFile) -> bytes:
    b = io.BytesIO()
    write_model(b, mf)
    return b.getvalue()


def save(path: str, mf: ModelFile) -> None:
    with open(path, "wb") as f:
        write_model(f, mf)


def read_model(buf: bytes | memoryview) -> ModelFile:
    mv = memoryview(buf)
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, mv, off)
        off += struct.calcsize(fmt)
        return v

    (magic,) = take("<I")
    if magic != GGML_FILE_MAGIC:
        raise ValueError("invalid model data (bad magic)")
    hp = dict(zip(HPARAM_NAMES, take("<11i")))
    n_mel, n_fft = take("<ii")
    filters = np.frombuffer(mv, dtype=np.float32, count=n_mel * n_fft, offset=off).reshape(n_mel, n_fft).copy()
    off += 4 * n_mel * n_fft
    (n_vocab,) = take("<i")
    vocab = []
    for _ in range(n_vocab):
        (ln,) = take("<I")
        vocab.append(bytes(mv[off:off + ln]))
        off += ln
    tensors = []
    while off < len(mv):
        n_dims, name_len, ttype = take("<iii")
        ne = take(f"<{n_dims}i")
        name = bytes(mv[off:off + name_len]).decode()
        off += name_len
        nrows = int(np.prod(ne[1:])) if n_dims > 1 else 1
        nbytes = gq.row_bytes(ttype, ne[0]) * nrows
        data = np.frombuffer(mv, dtype=np.uint8, count=nbytes, offset=off).copy()
        off += nbytes
        tensors.append(TensorRec(name, ttype, tuple(ne), data))
    return ModelFile(hp, filters, vocab, tensors)


def load(path: str) -> ModelFile:
    with open(path, "rb") as f:
        return read_model(f.read())


def build_model(hp: dict, filters: np.ndarray, weights_f32: dict, wtype: int) -> ModelFile:
    """weights_f32: name -> float32 array in numpy (outermost-first) shape; encodes each to the type the loader expects."""
    hp = dict(hp)
    hp["ftype"] = (gq.GGML_QNT_VERSION * 1000 if wtype in (gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0) else 0) + gq.GGML_FTYPE[wtype]
    shapes = expected_shapes(hp)
    recs = []
    for name in tensor_names(hp["n_audio_layer"]):
        ne = shapes[name]
        w = np.asarray(weights_f32[name], dtype=np.float32).reshape(tuple(reversed(ne)))
        tt = tensor_type_for(name, ne, wtype)
        if tt in (gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0):
            # the fork's pipeline is checkpoint -> F16 file -> ggml quantiser, so quantisation sees F16-rounded values
            w = w.astype(np.float16).astype(np.float32)
        recs.append(TensorRec(name, tt, ne, gq.quantize(w.reshape(-1, ne[0]), tt).reshape(-1)))
    return ModelFile(hp, np.asarray(filters, dtype=np.float32), [], recs)


def quantize_model(mf: ModelFile, wtype: int) -> ModelFile:
    """F16/F32 file -> Q8_0 / Q4_0 file, like examples/common-ggml.cpp:41-205 with to_quant={".*"}, to_skip=NO_QUANT."""
    hp = dict(mf.hparams)
    hp["ftype"] = gq.GGML_QNT_VERSION * 1000 + gq.GGML_FTYPE[wtype]
    recs = []
    for t in mf.tensors:
        if len(t.ne) == 2 and t.name not in NO_QUANT:
            w = gq.dequantize(t.data, t.ttype, t.ne[0])
            recs.append(TensorRec(t.name, wtype, t.ne, gq.quantize(w, wtype).reshape(-1)))
        else:
            recs.append(t)
    return ModelFile(hp, mf.filters, list(mf.vocab), recs)
