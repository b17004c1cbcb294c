"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, ggml CPU backend) in the build
container, where /root/reference exists.  Committed together with its outputs; the GPU box only reads the .npz.

    python tests/golden/make_golden.py [--full]

Inputs are seeded (qwen2_audio_whisper_ggml_b200.synth), so the model files themselves are not stored:
  tiny_<wtype>.npz   2-layer d=128 model, seed 1; 2 s chirp (seed 3): mel window + full embeddings
  full_<wtype>.npz   32-layer d=1280 model, seed 1234; 30 s chirp (seed 0): mel rows + a row subsample of the embeddings
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    pcm = synth.synth_pcm(32000, seed=3)
    for w in ("f16", "q8_0", "q4_0", "f32"):
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
