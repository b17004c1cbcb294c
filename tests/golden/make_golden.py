"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, ggml CPU backend) in the build
container, where /root/reference exists.  Committed together with its outputs; the GPU box only reads the .npz.

    python tests/golden/make_golden.py [--full]

Inputs are seeded (qwen2_audio_whisper_ggml_b200.synth), so the model files themselves are not stored:
  tiny_<wtype>.npz   2-layer d=128 model, seed 1; 2 s chirp (seed 3): mel window + full embeddings
  full_<wtype>.npz   32-layer d=1280 model, seed 1234; 30 s chirp (seed 0): mel rows + a row subsample of the embeddings
  fullset_<wtype>.npz  (--fullset) the same model on five clips -- chirp, ragged (n = 250 001), silence, low-noise tones and
                     samples/jfk (tests/golden/jfk_16k.wav, decoded once by decode_jfk.py): per clip 64 embedding rows, the L2 norm
                     of ALL 750 rows, the column means, and six mel rows
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from qwen2_audio_whisper_ggml_b200 import ggml_quant as gq, modelfile as mfm, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WT = {"f16": gq.GGML_TYPE_F16, "q8_0": gq.GGML_TYPE_Q8_0, "q4_0": gq.GGML_TYPE_Q4_0, "f32": gq.GGML_TYPE_F32}
FULL_ROWS = np.array([0, 1, 2, 100, 250, 374, 375, 500, 600, 700, 748, 749])
FULL_MEL_ROWS = np.array([0, 1, 17, 64, 100, 127])


def run(hp, seed, pcm, wname, threads):
    mf = synth.synth_model(hp, WT[wname], seed=seed)
    buf = mfm.to_bytes(mf)
    ctx = refbind.RefContext(buf)
    t0 = time.time()
    rc = ctx.full(pcm, n_threads=threads)
    dt = time.time() - t0
    assert rc == 0
    mel, emb = ctx.get_mel(), ctx.get_embeddings()
    ctx.free()
    return mel, emb, dt


FULLSET_ROWS = np.unique(np.concatenate([np.arange(0, 750, 12), [1, 2, 374, 375, 748, 749]]))[:64]


def read_wav_16k(path):
    """16 kHz mono int16 -> float32 / 32768, what read_wav does (examples/common.cpp:723-728)"""
    import wave
    with wave.open(path) as w:
        assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        return (np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def fullset_clips():
    return {
        "chirp": synth.synth_pcm(480000, seed=0),
        "ragged": synth.synth_pcm(250001, seed=5, kind="chirp"),
        "silence": synth.synth_pcm(480000, seed=0, kind="silence"),
        "tones": synth.synth_pcm(480000, seed=0, kind="tones"),
        "jfk": read_wav_16k(os.path.join(HERE, "jfk_16k.wav")),
    }


def make_fullset(threads):
    clips = fullset_clips()
    for w in ("f16", "q8_0", "q4_0"):
        buf = mfm.to_bytes(synth.synth_model(synth.FULL_HPARAMS, WT[w], seed=1234))
        ctx = refbind.RefContext(buf)
        out = {"row_idx": FULLSET_ROWS, "mel_row_idx": FULL_MEL_ROWS}
        for name, pcm in clips.items():
            t0 = time.time()
            assert ctx.full(pcm, n_threads=threads) == 0
            dt = time.time() - t0
            mel, emb = ctx.get_mel(), ctx.get_embeddings()
            out[f"{name}_n"] = pcm.size
            out[f"{name}_emb_rows"] = emb[FULLSET_ROWS]
            out[f"{name}_row_norm"] = np.linalg.norm(emb.astype(np.float64), axis=1).astype(np.float32)
            out[f"{name}_col_mean"] = emb.mean(axis=0)
            out[f"{name}_mel_rows"] = mel[FULL_MEL_ROWS][:, :3000]
            out[f"{name}_first20"] = emb.reshape(-1)[:20]          # what whisper_print_emb_enc prints (src:4191-4203)
            print(f"fullset {w} {name}: n {pcm.size} mel {mel.shape} emb {emb.shape} {dt:.1f}s", flush=True)
        ctx.free()
        np.savez_compressed(os.path.join(HERE, f"fullset_{w}.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--fullset", action="store_true")
    ap.add_argument("--no-tiny", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    if a.fullset:
        make_fullset(a.threads)
    pcm = synth.synth_pcm(32000, seed=3)
    for w in () if a.no_tiny else ("f16", "q8_0", "q4_0", "f32"):
        mel, emb, dt = run(synth.TINY_HPARAMS, 1, pcm, w, a.threads)
        np.savez_compressed(os.path.join(HERE, f"tiny_{w}.npz"), mel=mel[:, :220], emb=emb, n_len=mel.shape[1])
        print(f"tiny {w}: mel {mel.shape} emb {emb.shape} {dt:.2f}s")
    if a.full:
        pcm = synth.synth_pcm(480000, seed=0)
        for w in ("f16", "q8_0", "q4_0"):
            mel, emb, dt = run(synth.FULL_HPARAMS, 1234, pcm, w, a.threads)
            np.savez_compressed(os.path.join(HERE, f"full_{w}.npz"), mel_rows=mel[FULL_MEL_ROWS][:, :3000], mel_row_idx=FULL_MEL_ROWS,
                                emb_rows=emb[FULL_ROWS], emb_row_idx=FULL_ROWS, emb_mean=emb.mean(), emb_abs_mean=np.abs(emb).mean(),
                                emb_col_mean=emb.mean(axis=0), ref_seconds=dt, ref_threads=a.threads)
            print(f"full {w}: mel {mel.shape} emb {emb.shape} {dt:.1f}s on {a.threads} threads")


if __name__ == "__main__":
    main()
