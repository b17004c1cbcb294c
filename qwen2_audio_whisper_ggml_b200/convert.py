"""Checkpoint -> ggml `.bin` converter (the fork's models/convert-pt-to-ggml.py).

Accepts what the reference converter accepts -- a torch file holding {"dims": {...}, "model_state_dict": {...}}
(models/convert-pt-to-ggml.py:210-213) -- or a bare state dict whose keys are those of HF transformers'
`Qwen2AudioEncoder` (the `audio_tower` of Qwen2-Audio), which are exactly the tensor names the fork's loader expects
(src/qwen2-whisper.cpp:1591-1662).  Unlike the reference script, >= 2-D tensors are explicitly cast to F16 (the reference
stamps ftype = 1 but leaves tensors in their checkpoint dtype, :309-321), and Q8_0 / Q4_0 output is produced directly with
the bit-exact ggml block quantiser.

    python -m qwen2_audio_whisper_ggml_b200.convert checkpoint.pt out.bin [--wtype f16|q8_0|q4_0|f32] [--n-vocab 51866]
"""
from __future__ import annotations

import argparse

import numpy as np

from . import ggml_quant as gq
from . import modelfile as mfm
from . import synth

WTYPES = {"f32": gq.GGML_TYPE_F32, "f16": gq.GGML_TYPE_F16, "q8_0": gq.GGML_TYPE_Q8_0, "q4_0": gq.GGML_TYPE_Q4_0}


def hparams_from_state_dict(sd: dict, n_head: int | None = None, n_vocab: int = 51866) -> dict:
    D, T = [int(x) for x in sd["embed_positions.weight"].shape[::-1]]
    n_mels = int(sd["conv1.weight"].shape[1])
    n_layer = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    return dict(n_vocab=n_vocab, n_audio_ctx=T, n_audio_state=D, n_audio_head=n_head or D // 64, n_audio_layer=n_layer,
                n_text_ctx=448, n_text_state=D, n_text_head=n_head or D // 64, n_text_layer=0, n_mels=n_mels, ftype=1)


def from_state_dict(sd: dict, wtype: int = gq.GGML_TYPE_F16, hparams: dict | None = None, filters: np.ndarray | None = None) -> mfm.ModelFile:
    """sd: name -> array-like (torch tensors or numpy).  Keys outside the encoder (e.g. a leading 'audio_tower.') are handled."""
    clean = {}
    for k, v in sd.items():
        for pre in ("audio_tower.", "model.audio_tower.", "encoder."):
            if k.startswith(pre):
                k = k[len(pre):]
        a = v.detach().float().cpu().numpy() if hasattr(v, "detach") else np.asarray(v, dtype=np.float32)
        clean[k] = a
    hp = dict(hparams) if hparams else hparams_from_state_dict(clean)
    want = mfm.expected_shapes(hp)
    weights = {}
    for name, ne in want.items():
        if name not in clean:
            raise KeyError(f"checkpoint is missing tensor '{name}'")
        a = clean[name]
        if name in ("conv1.bias", "conv2.bias"):
            a = a.reshape(-1, 1)                       # the fork stores conv biases as 2-D [n_state, 1] (:1594, :1597)
        if tuple(a.shape) != tuple(reversed(ne)):
            raise ValueError(f"tensor '{name}' has shape {a.shape}, expected {tuple(reversed(ne))}")
        weights[name] = a.astype(np.float32)
    if filters is None:
        filters = synth.slaney_mel_filters(hp["n_mels"])    # == whisper's mel_filters.npz entry the reference script reads (:268-281)
    return mfm.build_model(hp, filters, weights, wtype)


def main(argv=None):
    import torch

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("checkpoint")
    ap.add_argument("output")
    ap.add_argument("--wtype", default="f16", choices=list(WTYPES))
    ap.add_argument("--n-vocab", type=int, default=51866)
    a = ap.parse_args(argv)
    ck = torch.load(a.checkpoint, map_location="cpu", weights_only=True)
    hp = None
    if isinstance(ck, dict) and "model_state_dict" in ck:
        dims = ck.get("dims", {})
        sd = ck["model_state_dict"]
        if dims:
            hp = dict(n_vocab=dims.get("n_vocab", a.n_vocab), n_audio_ctx=dims["n_audio_ctx"], n_audio_state=dims["n_audio_state"],
                      n_audio_head=dims["n_audio_head"], n_audio_layer=dims["n_audio_layer"], n_text_ctx=dims.get("n_text_ctx", 448),
                      n_text_state=dims.get("n_text_state", dims["n_audio_state"]), n_text_head=dims.get("n_text_head", dims["n_audio_head"]),
                      n_text_layer=0, n_mels=dims["n_mels"], ftype=1)
    else:
        sd = ck
    mf = from_state_dict(sd, WTYPES[a.wtype], hp)
    mfm.save(a.output, mf)
    print(f"wrote {a.output}: {len(mf.tensors)} tensors, wtype {a.wtype}, hparams {mf.hparams}")


if __name__ == "__main__":
    main()
