"""CPU tier: pin the oracle.  The numpy restatements (oracle/mel_np.py, oracle/encoder_np.py) are checked against
 (a) the golden vectors the unmodified reference produced (tests/golden, see make_golden.py) and
 (b) the compiled reference run live (oracle/_ref) on further inputs incl. the reference's edge cases."""
import os

import numpy as np
import pytest

from util import max_abs, rel_l2

from oracle import encoder_np, mel_np
from qwen2_audio_whisper_ggml_b200 import ggml_quant as gq, modelfile as mfm, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WT = {"f16": gq.GGML_TYPE_F16, "q8_0": gq.GGML_TYPE_Q8_0, "q4_0": gq.GGML_TYPE_Q4_0, "f32": gq.GGML_TYPE_F32}


@pytest.mark.parametrize("wname", ["f16", "q8_0", "q4_0", "f32"])
def test_restatement_vs_golden_tiny(wname):
    g = np.load(os.path.join(GOLD, f"tiny_{wname}.npz"))
    mf = synth.synth_model(synth.TINY_HPARAMS, WT[wname], seed=1)
    pcm = synth.synth_pcm(32000, seed=3)
    mel = mel_np.log_mel_spectrogram(pcm, mf.filters)
    assert mel.shape[1] == int(g["n_len"])
    assert max_abs(mel[:, :220], g["mel"]) < 5e-5
    emb = encoder_np.EncoderOracle(mf, "ggml").encode(mel_np.window(mel, 0, 100))
    # ggml-rounding mode tracks the CPU backend: 1.4e-4 for F16; for quantised weights the int8 rounding decisions of the
    # activations are chaotic under 1-ulp differences in summation order, which leaves ~3e-3 (still 2x closer than F32 math)
    lim_ggml = 1e-3 if wname in ("f16", "f32") else 5e-3
    assert rel_l2(emb, g["emb"]) < lim_ggml, rel_l2(emb, g["emb"])
    emb32 = encoder_np.EncoderOracle(mf, "f32").encode(mel_np.window(mel, 0, 100))
    lim = 2e-3 if wname in ("f16", "f32") else 3e-2
    assert rel_l2(emb32, g["emb"]) < lim


@pytest.mark.parametrize("kind,n", [("chirp", 480000), ("tones", 160000), ("silence", 16000), ("noise", 12345), ("chirp", 201), ("noise", 480160)])
def test_mel_restatement_vs_live_reference(ref, kind, n):
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=1)
    ctx = ref.RefContext(mfm.to_bytes(mf))
    pcm = synth.synth_pcm(n, seed=5, kind=kind)
    want = ctx.pcm_to_mel(pcm, n_threads=3)
    n_len, n_len_org, n_mel = ctx.mel_dims()
    assert (n_len, n_len_org) == mel_np.mel_dims(n) and n_mel == 128
    got = mel_np.log_mel_spectrogram(pcm, mf.filters)
    assert max_abs(got, want) < 5e-5, max_abs(got, want)
    ctx.free()


def test_reference_is_deterministic_across_thread_counts(ref):
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_Q8_0, seed=2)
    buf = mfm.to_bytes(mf)
    pcm = synth.synth_pcm(32000, seed=8)
    outs = []
    for th in (1, 3, 8):
        ctx = ref.RefContext(buf)
        assert ctx.full(pcm, n_threads=th) == 0
        outs.append(ctx.get_embeddings())
        ctx.free()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_gelu_table_restatement(ref):
    x = np.concatenate([np.linspace(-12, 12, 4001), [0.0, -0.0, 1e-4, 10.0, -10.0, 11.0]]).astype(np.float32)
    want = ref.ref_gelu(x)
    got = encoder_np.gelu_tanh(x, via_f16=True)
    # the table is built from F16 inputs and holds F16 outputs: agreement to one F16 ulp (2^-10 relative)
    assert np.all(np.abs(got - want) <= np.abs(want) * 2.0 ** -10 + 2e-6)   # + F16 subnormal spacing near zero
    assert np.mean(got == want) > 0.95


def test_encoder_restatement_offsets_vs_live_reference(ref):
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=7)
    ctx = ref.RefContext(mfm.to_bytes(mf))
    pcm = synth.synth_pcm(5 * 16000, seed=21, kind="noise")
    orc = encoder_np.EncoderOracle(mf, "ggml")
    mel = mel_np.log_mel_spectrogram(pcm, mf.filters)
    for off_ms in (0, 2500):
        assert ctx.full(pcm, offset_ms=off_ms) == 0
        emb = orc.encode(mel_np.window(mel, off_ms // 10, 100))
        assert rel_l2(emb, ctx.get_embeddings()) < 1e-3
    ctx.free()


def test_restatement_full_size_on_samples_jfk_vs_golden():
    """the oracle pinned at the size and on the input BASELINE configs[0] names: full model (32 layers, d = 1280), samples/jfk
    (tests/golden/jfk_16k.wav) -- mel of all five fullset clips and the encoder on the speech clip against the reference's output"""
    import wave
    g = np.load(os.path.join(GOLD, "fullset_f16.npz"))
    with wave.open(os.path.join(GOLD, "jfk_16k.wav")) as w:
        jfk = (np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    filters = synth.slaney_mel_filters(128)
    clips = {"chirp": synth.synth_pcm(480000, seed=0), "ragged": synth.synth_pcm(250001, seed=5, kind="chirp"),
             "silence": synth.synth_pcm(480000, seed=0, kind="silence"), "tones": synth.synth_pcm(480000, seed=0, kind="tones"), "jfk": jfk}
    for name, pcm in clips.items():
        mel = mel_np.log_mel_spectrogram(pcm, filters)
        assert max_abs(mel[g["mel_row_idx"]][:, :3000], g[f"{name}_mel_rows"]) < 5e-5, name
    mf = synth.synth_model(synth.FULL_HPARAMS, WT["f16"], seed=1234)
    mel = mel_np.log_mel_spectrogram(jfk, mf.filters)
    emb = encoder_np.EncoderOracle(mf, "ggml").encode(mel_np.window(mel, 0, 1500))
    rows, want = emb[g["row_idx"]], g["jfk_emb_rows"]
    assert rel_l2(rows, want) < 1e-3, rel_l2(rows, want)
    assert np.abs(np.linalg.norm(emb.astype(np.float64), axis=1) / g["jfk_row_norm"] - 1.0).max() < 2e-3
