"""-m gpu: PCM -> mel -> encoder parity through the reference-compatible C API (include/qwen2-whisper.h), against
 (a) golden vectors produced by the unmodified reference's ggml CPU backend (tests/golden, made by make_golden.py),
 (b) the compiled reference run live on the box's host cores (oracle/_ref), and
 (c) the numpy restatement (oracle/encoder_np.py) for the "dequantised-weight F32" check of the quantised paths.
Tolerances: util.TOL."""
import os

import numpy as np
import pytest

from util import TOL, max_abs, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from qwen2_audio_whisper_ggml_b200 import Context, ggml_quant as gq, modelfile as mfm, synth  # noqa: E402
from qwen2_audio_whisper_ggml_b200 import api  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WT = {"f16": gq.GGML_TYPE_F16, "q8_0": gq.GGML_TYPE_Q8_0, "q4_0": gq.GGML_TYPE_Q4_0, "f32": gq.GGML_TYPE_F32}


@pytest.fixture(scope="module", autouse=True)
def quiet_log():
    api.log_set(lambda lvl, txt: None)
    yield
    api.log_set(None)


def tiny_ctx(wname, seed=1):
    buf = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, WT[wname], seed=seed))
    return Context.init_from_buffer(buf), buf


@pytest.mark.parametrize("wname", ["f16", "q8_0", "q4_0", "f32"])
def test_tiny_full_vs_golden(wname):
    g = np.load(os.path.join(GOLD, f"tiny_{wname}.npz"))
    ctx, _ = tiny_ctx(wname)
    pcm = synth.synth_pcm(32000, seed=3)
    assert ctx.full(pcm) == 0
    assert ctx.mel_dims()[0] == int(g["n_len"])
    assert ctx.n_len() == 1 + (32000 + 200 - 400) // 160          # whisper_n_len == n_len_org (src:3440-3446)
    mel = ctx.get_mel()
    assert max_abs(mel[:, :220], g["mel"]) < TOL["mel"]["max_abs"]
    emb = ctx.get_embeddings()
    assert emb.shape == (1,) + g["emb"].shape
    tol = TOL["f16" if wname in ("f16", "f32") else wname]
    assert rel_l2(emb[0], g["emb"]) < tol["rel_l2"], rel_l2(emb[0], g["emb"])
    assert max_abs(emb[0], g["emb"]) < tol["max_abs"]
    ctx.free()


@pytest.mark.parametrize("wname", ["q8_0", "q4_0"])
def test_tiny_quant_vs_f32_restatement(wname):
    """block decode + GEMM are exact: compare with plain F32 math on the dequantised weights (no activation quantisation)"""
    from oracle import encoder_np, mel_np
    ctx, buf = tiny_ctx(wname)
    pcm = synth.synth_pcm(32000, seed=3)
    assert ctx.full(pcm) == 0
    mf = mfm.read_model(buf)
    want = encoder_np.EncoderOracle(mf, "f32").encode(mel_np.window(mel_np.log_mel_spectrogram(pcm, mf.filters), 0, 100))
    emb = ctx.get_embeddings()[0]
    t = TOL["quant_vs_f32_restatement"]
    assert rel_l2(emb, want) < t["rel_l2"], rel_l2(emb, want)
    assert max_abs(emb, want) < t["max_abs"]
    ctx.free()


def test_tiny_live_reference_offsets_and_set_mel(ref):
    """whisper_full with offset_ms, whisper_set_mel + whisper_encode, against the reference run live"""
    ctx, buf = tiny_ctx("f16", seed=7)
    rctx = ref.RefContext(buf)
    pcm = synth.synth_pcm(5 * 16000, seed=21, kind="noise")
    for off_ms in (0, 1000, 2500):
        assert ctx.full(pcm, offset_ms=off_ms) == 0
        assert rctx.full(pcm, offset_ms=off_ms) == 0
        a, b = ctx.get_embeddings()[0], rctx.get_embeddings()
        assert rel_l2(a, b) < TOL["f16"]["rel_l2"], (off_ms, rel_l2(a, b))
    mel = rctx.get_mel()
    assert max_abs(ctx.get_mel(), mel) < TOL["mel"]["max_abs"]
    # caller-provided mel, window running past n_len is zero-filled (src:2274-2283)
    short = np.ascontiguousarray(mel[:, :150])
    assert ctx.set_mel(short, 150, 128) == 0 and rctx.set_mel(short) == 0
    assert ctx.n_len() == 150
    assert ctx.encode(40) == 0
    assert rctx.full(None, offset_ms=400) == 0
    assert rel_l2(ctx.get_embeddings()[0], rctx.get_embeddings()) < TOL["f16"]["rel_l2"]
    assert ctx.set_mel(short, 150, 80) == -1          # wrong n_mel is rejected like the reference (:3287)
    ctx.free()
    rctx.free()


def test_short_audio_returns_zero_and_does_nothing(ref):
    ctx, buf = tiny_ctx("f16")
    p = api.wlib().whisper_full_default_params()
    p.offset_ms = 0
    p.duration_ms = 500           # < 1000 ms -> warning + 0 (:2362-2365)
    assert ctx.full(synth.synth_pcm(32000, seed=1), params=p) == 0
    assert ctx.embd_dims()[0] == 0
    # a 12345-sample clip has n_len_org = 76 < 100 frames: the reference returns 0 without running the encoder, and so do we
    short = synth.synth_pcm(12345, seed=2, kind="noise")
    rctx = ref.RefContext(buf)
    assert rctx.full(short) == 0 and rctx.timings()["n_encode"] == 0      # returned 0 but never ran the encoder
    assert ctx.full(short) == 0 and ctx.embd_dims()[0] == 0
    assert ctx.n_len() == rctx.mel_dims()[1] == 76
    rctx.free()
    ctx.free()


def test_batch_equals_single_windows_bitwise():
    """data-parallel contract: window b of a batch == the same window alone, bit for bit; ragged lengths; chunked micro-batches"""
    ctx, _ = tiny_ctx("f16")
    win = 200 * 160
    B = 5
    pcm = np.zeros((B, win), dtype=np.float32)
    ns = np.array([win, win // 2, win, 16000, win - 7], dtype=np.int32)
    for b in range(B):
        pcm[b, :ns[b]] = synth.synth_pcm(int(ns[b]), seed=40 + b, kind="chirp" if b % 2 else "noise")
    assert ctx.set_max_batch(2) == 0
    out = ctx.encode_batch(pcm, ns)
    assert ctx.set_max_batch(8) == 0
    out8 = ctx.encode_batch(pcm, ns)
    assert np.array_equal(out, out8)
    for b in range(B):
        # whisper_pcm_to_mel + whisper_encode (whisper_full would skip the clips under 1 s: n_len_org < 100, src:2362)
        assert ctx.pcm_to_mel(pcm[b, :ns[b]]) == 0 and ctx.encode(0) == 0
        single = ctx.get_embeddings()[0]
        # the API mel covers n + 30 s of padding, the batch API one window: same frames, same max
        assert rel_l2(out[b], single) < 1e-6, (b, rel_l2(out[b], single))
    ctx.free()


def test_loader_error_paths():
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=1)
    good = mfm.to_bytes(mf)
    bad_magic = b"\x00\x00\x00\x00" + good[4:]
    with pytest.raises(Exception):
        Context.init_from_buffer(bad_magic)
    with pytest.raises(Exception):                                  # truncated: not all tensors loaded (:1861)
        Context.init_from_buffer(mfm.to_bytes(mfm.ModelFile(mf.hparams, mf.filters, [], mf.tensors[:-1])))
    wrong = list(mf.tensors)
    t = wrong[0]                                                    # embed_positions.weight [D, T] -> [T, D]
    wrong[0] = mfm.TensorRec(t.name, t.ttype, t.ne[::-1], t.data)
    with pytest.raises(Exception):                                  # wrong shape (:1821)
        Context.init_from_buffer(mfm.to_bytes(mfm.ModelFile(mf.hparams, mf.filters, [], wrong)))
    unk = list(mf.tensors) + [mfm.TensorRec("decoder.blocks.0.attn.weight", 0, (4,), np.zeros(16, np.uint8))]
    with pytest.raises(Exception):                                  # unknown tensor (:1807)
        Context.init_from_buffer(mfm.to_bytes(mfm.ModelFile(mf.hparams, mf.filters, [], unk)))
    p = api.default_context_params()
    p.flash_attn = True
    with pytest.raises(Exception):
        Context.init_from_buffer(good, p)
    p = api.default_context_params()
    p.use_gpu = False
    with pytest.raises(Exception):
        Context.init_from_buffer(good, p)
    Context.init_from_buffer(good).free()


@pytest.mark.parametrize("wname", ["f16", "q8_0", "q4_0"])
def test_full_size_vs_golden(wname):
    """BASELINE config 0 shape: 32 layers, d=1280, 20 heads, 128 mels, one 30 s window; golden from the reference's CPU backend"""
    path = os.path.join(GOLD, f"full_{wname}.npz")
    if not os.path.exists(path):
        pytest.skip("full-size golden not generated")
    g = np.load(path)
    buf = mfm.to_bytes(synth.synth_model(synth.FULL_HPARAMS, WT[wname], seed=1234))
    ctx = Context.init_from_buffer(buf)
    del buf
    pcm = synth.synth_pcm(480000, seed=0)
    assert ctx.full(pcm) == 0
    mel = ctx.get_mel()
    assert mel.shape == (128, 6000)
    assert max_abs(mel[g["mel_row_idx"]][:, :3000], g["mel_rows"]) < TOL["mel"]["max_abs"]
    emb = ctx.get_embeddings()[0]
    assert emb.shape == (750, 1280)
    tol = TOL[wname]
    rows = emb[g["emb_row_idx"]]
    assert rel_l2(rows, g["emb_rows"]) < tol["rel_l2"], rel_l2(rows, g["emb_rows"])
    assert max_abs(rows, g["emb_rows"]) < tol["max_abs"], max_abs(rows, g["emb_rows"])
    assert max_abs(emb.mean(axis=0), g["emb_col_mean"]) < tol["max_abs"]
    # batched path on the same clip (3 copies + a silent window) == single-window path
    ctx.set_max_batch(4)
    win = np.stack([pcm, pcm, np.zeros_like(pcm), pcm])
    out = ctx.encode_batch(win)
    # the single window runs the small-M split-K residual epilogue (order of the F32 adds differs), the 4-window batch does not: a
    # one-ulp difference in the residual stream flips F16 roundings downstream and settles at the model's F16 noise floor (~3e-4,
    # the same distance either result has from the reference)
    assert rel_l2(out[0], emb) < 1e-3 and np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[3])
    assert np.isfinite(out[2]).all()
    ctx.free()


def test_long_audio_chunked_equals_per_window_reference(ref):
    """BASELINE config 5 semantics: long PCM cut into windows, per-window mel max; each window == the reference on that window"""
    ctx, buf = tiny_ctx("f16", seed=3)
    rctx = ref.RefContext(buf)
    win = 200 * 160
    pcm = synth.synth_pcm(3 * win + 20000, seed=77, kind="noise")   # ragged tail >= 1 s (the reference skips shorter clips)
    out = ctx.encode_long(pcm)
    assert out.shape == (4, 50, 128)
    for w in range(4):
        seg = pcm[w * win:(w + 1) * win]
        assert rctx.full(seg) == 0
        assert rel_l2(out[w], rctx.get_embeddings()) < TOL["f16"]["rel_l2"], w
    ctx.free()
    rctx.free()


def test_second_state_shares_the_model():
    """several states over one read-only model (src/qwen2-whisper.cpp:769-770), whisper_init_state / whisper_free_state"""
    import ctypes as C
    ctx, _ = tiny_ctx("f16")
    w = api.wlib()
    st = w.whisper_init_state(ctx._h)
    assert st
    a = synth.synth_pcm(32000, seed=5)
    b = synth.synth_pcm(32000, seed=6, kind="noise")
    p = w.whisper_full_default_params()
    assert ctx.full(a) == 0
    assert w.whisper_full_with_state(ctx._h, st, p, b.ctypes.data, b.size) == 0
    e_default = ctx.get_embeddings()[0]
    e_state = np.empty_like(e_default)
    assert w.whisper_get_embeddings_from_state(st, e_state.ctypes.data, e_state.size) == 0
    assert ctx.full(b) == 0
    assert np.array_equal(ctx.get_embeddings()[0], e_state) and not np.array_equal(e_default, e_state)
    assert w.whisper_n_len_from_state(st) == ctx.n_len()
    w.whisper_free_state(st)
    ctx.free()


def test_cli_matches_reference_printout(ref, tmp_path, capfd):
    """q2w-main (examples/main equivalent): WAV in, ' %.3f' x 20 out -- the reference's only user-visible check"""
    import subprocess
    import wave
    exe = os.path.join(os.path.dirname(GOLD), "..", "qwen2_audio_whisper_ggml_b200", "q2w-main")
    exe = os.path.abspath(exe)
    if not os.path.exists(exe):
        pytest.skip("q2w-main not built")
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=1)
    model = tmp_path / "tiny-f16.bin"
    mfm.save(str(model), mf)
    pcm = synth.synth_pcm(40000, seed=9)
    wav = tmp_path / "a.wav"
    with wave.open(str(wav), "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(16000)
        wf.writeframes(np.round(pcm * 32768.0).astype(np.int16).tobytes())
    r = subprocess.run([exe, "-m", str(model), "-f", str(wav), "-n", "2", "-np"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 2 and lines[0] == lines[1]
    got = np.array([float(x) for x in lines[0].split()])
    rctx = ref.RefContext(mfm.to_bytes(mf))
    assert rctx.full(pcm) == 0
    want = rctx.get_embeddings().reshape(-1)[:20]
    assert got.shape == (20,) and np.abs(got - want).max() < 5e-3      # %.3f rounding + F16 tolerance
    rctx.free()


def test_hf_qwen2_audio_encoder_checkpoint_through_engine():
    """independent oracle: HF transformers' Qwen2AudioEncoder (tanh GELU) -> converter -> this library == HF's own forward"""
    pytest.importorskip("transformers")
    from test_convert_cpu import tiny_hf_encoder
    from oracle import mel_np
    from qwen2_audio_whisper_ggml_b200 import convert
    enc = tiny_hf_encoder(1)
    sd = dict(enc.state_dict())
    hp = convert.hparams_from_state_dict(sd, n_head=2, n_vocab=64)
    pcm = synth.synth_pcm(32000, seed=4)
    for wt, tol in ((gq.GGML_TYPE_F16, 2e-3), (gq.GGML_TYPE_F32, 2e-3), (gq.GGML_TYPE_Q8_0, 3e-2)):
        mf = convert.from_state_dict(sd, wt, hp)
        ctx = Context.init_from_buffer(mfm.to_bytes(mf))
        assert ctx.full(pcm) == 0
        win = mel_np.window(ctx.get_mel(), 0, 100)
        with torch.no_grad():
            want = enc(torch.from_numpy(win)[None]).last_hidden_state[0].numpy()
        got = ctx.get_embeddings()[0]
        assert rel_l2(got, want) < tol, (gq.TYPE_NAMES[wt], rel_l2(got, want))
        ctx.free()


def test_init_with_custom_loader_callbacks():
    """whisper_init_with_params(loader): read / eof / close are invoked synchronously and close is always called (src:3111-3137)"""
    import ctypes as C
    w = api.wlib()
    buf = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_Q4_0, seed=1))
    state = {"off": 0, "closed": 0, "hit_end": False}

    class Loader(C.Structure):
        _fields_ = [("context", C.c_void_p), ("read", C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t)),
                    ("eof", C.CFUNCTYPE(C.c_bool, C.c_void_p)), ("close", C.CFUNCTYPE(None, C.c_void_p))]

    def rd(ctx, out, n):
        k = min(n, len(buf) - state["off"])
        if k < n:
            state["hit_end"] = True
        C.memmove(out, buf[state["off"]:state["off"] + k], k)
        state["off"] += k
        return k

    ld = Loader(None, Loader._fields_[1][1](rd), Loader._fields_[2][1](lambda c: state["hit_end"]),
                Loader._fields_[3][1](lambda c: state.__setitem__("closed", state["closed"] + 1)))
    w.whisper_init_with_params.restype = C.c_void_p
    w.whisper_init_with_params.argtypes = [C.POINTER(Loader), api.ContextParams]
    h = w.whisper_init_with_params(C.byref(ld), api.default_context_params())
    assert h and state["closed"] == 1 and state["off"] == len(buf)
    ctx = Context(h)
    assert ctx.model_n("ftype") == 2 and ctx.model_n("n_audio_layer") == 2
    g = np.load(os.path.join(GOLD, "tiny_q4_0.npz"))
    assert ctx.full(synth.synth_pcm(32000, seed=3)) == 0
    assert rel_l2(ctx.get_embeddings()[0], g["emb"]) < TOL["q4_0"]["rel_l2"]
    ctx.free()
    # a loader that fails mid-way still gets closed, and init returns NULL
    state.update(off=0, closed=0, hit_end=False)
    short = buf[: len(buf) // 2]
    def rd2(ctx, out, n):
        k = min(n, len(short) - state["off"])
        if k < n:
            state["hit_end"] = True
        C.memmove(out, short[state["off"]:state["off"] + k], k)
        state["off"] += k
        return k
    ld2 = Loader(None, Loader._fields_[1][1](rd2), ld.eof, ld.close)
    assert not w.whisper_init_with_params(C.byref(ld2), api.default_context_params())
    assert state["closed"] == 1


def test_whole_file_streaming_matches_reference_offsets(ref):
    """one globally normalised mel over a long clip, all sliding windows as one batch ==
    the reference's whisper_full(n_samples > 0) followed by whisper_full(NULL, 0, offset_ms = k * hop)"""
    ctx, buf = tiny_ctx("f16", seed=5)
    rctx = ref.RefContext(buf)
    win_frames = 200
    pcm = synth.synth_pcm(7 * 16000 + 777, seed=31, kind="chirp")          # 7.05 s -> 4 windows of 2 s, the last ragged
    pcm[5 * 16000:] *= 0.01                                                 # quiet tail: global vs per-window normalisation differ here
    out = ctx.encode_stream(pcm)
    assert out.shape[0] == 4
    assert rctx.full(pcm) == 0                                              # computes the mel once + window 0
    for k in range(4):
        if k:
            assert rctx.full(None, offset_ms=k * win_frames * 10) == 0     # mel NOT recomputed (n_samples == 0)
        assert rel_l2(out[k], rctx.get_embeddings()) < TOL["f16"]["rel_l2"], k
    # per-window normalisation (encode_long) must differ on the quiet tail, or this test proves nothing
    per_window = ctx.encode_long(pcm)
    assert rel_l2(per_window[3], out[3]) > 1e-2
    # half-window hop: 50 % overlap
    out2 = ctx.encode_stream(pcm, hop_frames=100)
    assert out2.shape[0] == 8 and np.array_equal(out2[0], out[0]) and np.array_equal(out2[2], out[1])
    ctx.free()
    rctx.free()


def test_async_batches_pipeline_and_match_the_synchronous_call():
    """whisper_encode_batch_async / _wait: two batches in flight, a third submit first retires the oldest; results are bit-identical
    to whisper_encode_batch, and the accessors follow the last batch waited for"""
    import torch
    ctx, _ = tiny_ctx("f16", seed=9)
    win = 2 * ctx.model_n("n_audio_ctx") * 160
    n_out, n_state = ctx.model_n("n_audio_ctx") // 2, ctx.model_n("n_audio_state")
    B = 20                                                                   # >= 16: the host path cuts it into two micro-batches
    ctx.set_max_batch(B)
    batches = [np.stack([synth.synth_pcm(win, seed=100 * k + w, kind="chirp") for w in range(B)]) for k in range(3)]
    want = [ctx.encode_batch(b).copy() for b in batches]
    ins = [torch.from_numpy(b).pin_memory() for b in batches]
    outs = [torch.full((B, n_out, n_state), float("nan")).pin_memory() for _ in range(3)]
    t0 = ctx.encode_batch_async(ins[0].numpy(), outs[0].numpy())
    t1 = ctx.encode_batch_async(ins[1].numpy(), outs[1].numpy())
    t2 = ctx.encode_batch_async(ins[2].numpy(), outs[2].numpy())              # retires t0 internally
    assert (t0, t1, t2) == (0, 1, 2)
    with pytest.raises(Exception):
        ctx.wait(t0)                                                         # already retired
    ctx.wait(t1)
    assert np.array_equal(ctx.get_embeddings(), want[1])                     # accessors: the batch just waited for
    ctx.wait(t2)
    for k in range(3):
        assert np.array_equal(outs[k].numpy(), want[k]), k
    assert np.array_equal(ctx.get_embeddings(), want[2])
    with pytest.raises(Exception):
        ctx.wait(7)
    # a synchronous call drains whatever is in flight and still works
    t3 = ctx.encode_batch_async(ins[0].numpy(), outs[0].numpy())
    again = ctx.encode_batch(batches[1])
    assert np.array_equal(again, want[1]) and np.array_equal(outs[0].numpy(), want[0])
    with pytest.raises(Exception):
        ctx.wait(t3)                                                         # drained by the synchronous call
    ctx.free()


def test_full_size_batch64_is_permutation_equivariant_and_shardable():
    """BASELINE configs[1] scale (full model, 64 x 30 s windows): size-independent properties of the data-parallel path --
    permuting the windows permutes the embeddings bit for bit, and any contiguous shard of the batch (what another rank would
    be given) reproduces its slice bit for bit; every output row is a LayerNorm output (finite, unit variance against gamma = 1 +- small)"""
    buf = mfm.to_bytes(synth.synth_model(synth.FULL_HPARAMS, WT["f16"], seed=1234))
    ctx = Context.init_from_buffer(buf)
    del buf
    B = 64
    ctx.set_max_batch(B)
    rng = np.random.default_rng(5)
    base = [synth.synth_pcm(480000, seed=200 + k, kind="chirp" if k % 2 else "noise") for k in range(8)]
    win = np.stack([np.roll(base[w % 8], 997 * w) * (0.25 + 0.75 * rng.random()) for w in range(B)]).astype(np.float32)
    ns = np.full(B, 480000, dtype=np.int32)
    ns[7], ns[40], ns[63] = 16000, 250001, 479999                      # ragged windows inside the batch
    out = ctx.encode_batch(win, ns)
    assert out.shape == (B, 750, 1280) and np.isfinite(out).all()
    perm = rng.permutation(B)
    out_p = ctx.encode_batch(np.ascontiguousarray(win[perm]), ns[perm])
    assert np.array_equal(out_p, out[perm])
    from qwen2_audio_whisper_ggml_b200.parallel import shard_bounds
    for rank in (0, 2):                                                 # 3-way sharding: 22 / 21 / 21 windows
        lo, hi = shard_bounds(B, rank, 3)
        assert np.array_equal(ctx.encode_batch(np.ascontiguousarray(win[lo:hi]), ns[lo:hi]), out[lo:hi]), rank
    assert len({out[w].tobytes() for w in range(B)}) == B              # no two windows collapsed onto each other
    v = out.reshape(-1, 1280).var(axis=1)
    assert 0.5 < float(v.min()) and float(v.max()) < 2.0, (v.min(), v.max())
    ctx.free()


# ------------------------------------------------------------------------------------------------ round 2
def test_stage_taps_tiny_vs_restatement():
    """conv stem (+ positional embedding) and the residual stream after each encoder block against the numpy restatement in its
    reference-rounding mode, so a drift can be localised to a stage"""
    from oracle import encoder_np, mel_np
    ctx, buf = tiny_ctx("f16", seed=11)
    mf = mfm.read_model(buf)
    orc = encoder_np.EncoderOracle(mf, "ggml")
    pcm = synth.synth_pcm(32000, seed=8)
    win = mel_np.window(mel_np.log_mel_spectrogram(pcm, mf.filters), 0, synth.TINY_HPARAMS["n_audio_ctx"])
    assert ctx.pcm_to_mel(pcm) == 0
    for k in range(synth.TINY_HPARAMS["n_audio_layer"] + 1):
        ctx.debug_forward_layers(k)
        assert ctx.encode(0) == 0
        got = ctx.debug_residual(0)
        want = orc.encode(win, return_pre_pool=True, n_layers=k)
        assert rel_l2(got, want) < 1e-3, (k, rel_l2(got, want))
    ctx.debug_forward_layers(-1)
    assert ctx.encode(0) == 0
    assert rel_l2(ctx.get_embeddings()[0], orc.encode(win)) < TOL["f16"]["rel_l2"]
    ctx.free()


def test_set_max_batch_keeps_mel_and_embeddings():
    """whisper_set_max_batch resizes scratch in place: the state's mel and last embeddings survive (round-1 dropped the state)"""
    ctx, _ = tiny_ctx("f16")
    pcm = synth.synth_pcm(32000, seed=3)
    assert ctx.full(pcm) == 0
    mel, emb = ctx.get_mel(), ctx.get_embeddings()
    assert ctx.set_max_batch(7) == 0
    assert np.array_equal(ctx.get_mel(), mel) and np.array_equal(ctx.get_embeddings(), emb)
    assert ctx.encode(0) == 0 and np.array_equal(ctx.get_embeddings(), emb)          # the kept mel still encodes to the same result
    assert ctx.set_max_batch(1) == 0 and ctx.encode(0) == 0 and np.array_equal(ctx.get_embeddings(), emb)
    ctx.free()


def test_truncated_file_is_rejected_wherever_it_is_cut():
    """a short read anywhere (filterbank, tensor name, inside the LAST tensor's payload) fails the load instead of uploading stale bytes"""
    good = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=1))
    Context.init_from_buffer(good).free()
    for cut in (len(good) - 1, len(good) - 100, len(good) // 2, 60, 4 + 44 + 8 + 1000):
        with pytest.raises(Exception):
            Context.init_from_buffer(good[:cut])


def test_deprecated_init_spellings(tmp_path):
    """whisper_init_from_file / _from_buffer (+ _no_state) still resolve and behave like the _with_params forms (src:3184-3206)"""
    import ctypes as C
    w = api.wlib()
    mf = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=1)
    path = tmp_path / "tiny.bin"
    mfm.save(str(path), mf)
    buf = mfm.to_bytes(mf)
    raw = (C.c_char * len(buf)).from_buffer_copy(buf)
    pcm = synth.synth_pcm(32000, seed=3)
    g = np.load(os.path.join(GOLD, "tiny_f16.npz"))
    for h in (w.whisper_init_from_file(str(path).encode()), w.whisper_init_from_buffer(C.cast(raw, C.c_void_p), len(buf))):
        ctx = Context(h)
        assert ctx.full(pcm) == 0 and rel_l2(ctx.get_embeddings()[0], g["emb"]) < TOL["f16"]["rel_l2"]
        ctx.free()
    for h in (w.whisper_init_from_file_no_state(str(path).encode()), w.whisper_init_from_buffer_no_state(C.cast(raw, C.c_void_p), len(buf))):
        ctx = Context(h)
        assert ctx.full(pcm) == -1            # no default state: the reference dereferences NULL here, we return an error
        st = w.whisper_init_state(ctx._h)
        assert st and w.whisper_full_with_state(ctx._h, st, w.whisper_full_default_params(), pcm.ctypes.data, pcm.size) == 0
        w.whisper_free_state(st)
        ctx.free()


def test_huge_offset_is_clamped_like_the_reference():
    """whisper_encode(offset near INT_MAX): i0 = min(offset, n_len) (src:2274), i.e. an all-zero window -- no overflow, no fault"""
    ctx, _ = tiny_ctx("f16")
    assert ctx.pcm_to_mel(synth.synth_pcm(32000, seed=3)) == 0
    n_len = ctx.mel_dims()[0]
    assert ctx.encode(n_len) == 0
    at_end = ctx.get_embeddings()
    assert ctx.encode(2 ** 31 - 1) == 0
    assert np.array_equal(ctx.get_embeddings(), at_end) and np.isfinite(at_end).all()
    ctx.free()


def _multi_case(devices):
    buf = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, WT["f16"], seed=13))
    win = 2 * synth.TINY_HPARAMS["n_audio_ctx"] * 160
    B = 11
    pcm = np.stack([synth.synth_pcm(win, seed=300 + w, kind="chirp" if w % 3 else "noise") for w in range(B)])
    ns = np.full(B, win, dtype=np.int32)
    ns[2], ns[9] = 16000, win - 5
    single = Context.init_from_buffer(buf)
    single.set_max_batch(4)
    want = single.encode_batch(pcm, ns)
    single.free()
    multi = Context.init_from_buffer(buf, devices=devices)
    assert multi.n_devices() == len(devices) and multi.devices() == list(devices)
    multi.set_max_batch(4)
    out = multi.encode_batch(pcm, ns)                          # whisper_encode_batch shards over the replicas
    assert np.array_equal(out, want)
    for k in sorted(set(devices)):
        out2 = multi.encode_batch_multi(pcm, ns, gather_device=k)
        assert np.array_equal(out2, want)
        assert np.array_equal(multi.gathered(B), want)         # gathered on device k, ordered by window index
        assert multi.gathered_device_ptr()
    # asynchronous batches over all replicas: two in flight, a third submit retires the oldest; same bits as the synchronous call
    import torch as _t
    ins = [_t.from_numpy(pcm).pin_memory() for _ in range(3)]
    outs = [_t.full((B,) + want.shape[1:], float("nan")).pin_memory() for _ in range(3)]
    tk = [multi.encode_batch_async(ins[k].numpy(), outs[k].numpy(), ns) for k in range(3)]
    assert tk == [0, 1, 2]
    with pytest.raises(Exception):
        multi.wait(tk[0])                                      # already retired by the third submit
    multi.wait(tk[1])
    multi.wait(tk[2])
    for k in range(3):
        assert np.array_equal(outs[k].numpy(), want), k
    t3 = multi.encode_batch_async(ins[0].numpy(), outs[0].numpy(), ns)
    assert np.array_equal(multi.encode_batch(pcm, ns), want)   # a synchronous call drains what is in flight
    with pytest.raises(Exception):
        multi.wait(t3)
    small = multi.encode_batch_multi(pcm[:1], ns[:1], gather_device=devices[-1])   # fewer windows than replicas: empty shards
    assert np.array_equal(small, want[:1]) and np.array_equal(multi.gathered(1), want[:1])
    # the single-window API keeps working on replica 0
    assert multi.full(pcm[0]) == 0 and rel_l2(multi.get_embeddings()[0], want[0]) < 1e-6
    with pytest.raises(Exception):
        multi.encode_batch_multi(pcm, ns, gather_device=63)    # not one of this context's devices
    multi.free()


def test_multi_replica_sharding_and_gather_on_one_device():
    """the one-process multi-device path (q2w_multi_*: replicas, worker threads, window w -> replica floor(w G / B), gather by peer
    copies) with three replicas on device 0: == the single-replica result bit for bit, in caller order"""
    _multi_case([0, 0, 0])


def test_multi_device_in_process_equals_single_device():
    """two real devices in ONE process (per-device function attributes, SM counts, streams): bit-identical to one device"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _multi_case([0, 1])
    p = api.default_context_params()
    p.gpu_device = -1                                          # -1 = every visible device
    ctx = Context.init_from_buffer(mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, WT["f16"], seed=13)), p)
    assert ctx.n_devices() == torch.cuda.device_count()
    ctx.free()
    p.gpu_device = 1                                           # a context on device 1 alone
    ctx = Context.init_from_buffer(mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, WT["f16"], seed=1)), p)
    g = np.load(os.path.join(GOLD, "tiny_f16.npz"))
    assert ctx.full(synth.synth_pcm(32000, seed=3)) == 0 and rel_l2(ctx.get_embeddings()[0], g["emb"]) < TOL["f16"]["rel_l2"]
    ctx.free()


def _fullset_clips():
    import wave
    with wave.open(os.path.join(GOLD, "jfk_16k.wav")) as w:
        jfk = (np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    return {"chirp": synth.synth_pcm(480000, seed=0), "ragged": synth.synth_pcm(250001, seed=5, kind="chirp"),
            "silence": synth.synth_pcm(480000, seed=0, kind="silence"), "tones": synth.synth_pcm(480000, seed=0, kind="tones"), "jfk": jfk}


@pytest.mark.parametrize("wname", ["f16", "q8_0", "q4_0"])
def test_full_size_five_clips_vs_golden(wname):
    """full-size model on five clips -- chirp, ragged (n = 250 001), silence, low-noise tones, and real speech (samples/jfk, BASELINE
    configs[0]) -- against the reference's ggml CPU backend: 64 embedding rows each, the L2 norm of ALL 750 rows, column means, mel
    rows, and the 20 values whisper_print_emb_enc prints; then the five clips again as ONE batch through whisper_encode_batch"""
    path = os.path.join(GOLD, f"fullset_{wname}.npz")
    if not os.path.exists(path):
        pytest.skip("fullset golden not generated")
    g = np.load(path)
    ctx = Context.init_from_buffer(mfm.to_bytes(synth.synth_model(synth.FULL_HPARAMS, WT[wname], seed=1234)))
    tol = TOL[wname]
    clips = _fullset_clips()
    singles = {}
    for name, pcm in clips.items():
        assert int(g[f"{name}_n"]) == pcm.size
        assert ctx.full(pcm) == 0, name
        mel = ctx.get_mel()
        assert max_abs(mel[g["mel_row_idx"]][:, :3000], g[f"{name}_mel_rows"]) < TOL["mel"]["max_abs"], name
        emb = ctx.get_embeddings()[0]
        singles[name] = emb
        rows, want = emb[g["row_idx"]], g[f"{name}_emb_rows"]
        assert rel_l2(rows, want) < tol["rel_l2"], (name, rel_l2(rows, want))
        assert max_abs(rows, want) < tol["max_abs"], (name, max_abs(rows, want))
        norms = np.linalg.norm(emb.astype(np.float64), axis=1)
        assert np.abs(norms / g[f"{name}_row_norm"] - 1.0).max() < 5 * tol["rel_l2"], (name, np.abs(norms / g[f"{name}_row_norm"] - 1.0).max())
        assert max_abs(emb.mean(axis=0), g[f"{name}_col_mean"]) < tol["max_abs"], name
        assert max_abs(emb.reshape(-1)[:20], g[f"{name}_first20"]) < tol["max_abs"], name
    # the same five clips as one ragged batch (chunk-then-mel semantics == whisper_full per clip: every clip is <= one window)
    ctx.set_max_batch(5)
    win = np.zeros((5, 480000), dtype=np.float32)
    ns = np.zeros(5, dtype=np.int32)
    for i, (name, pcm) in enumerate(clips.items()):
        win[i, :pcm.size] = pcm
        ns[i] = pcm.size
    out = ctx.encode_batch(win, ns)
    for i, name in enumerate(clips):
        rows, want = out[i][g["row_idx"]], g[f"{name}_emb_rows"]
        assert rel_l2(rows, want) < tol["rel_l2"], (name, rel_l2(rows, want))
        assert rel_l2(out[i], singles[name]) < 1e-3, name          # split-K single window vs batched single pass: F16 noise floor
    ctx.free()


@pytest.mark.parametrize("wname", ["f16", "q8_0", "q4_0"])
def test_full_size_stage_taps_and_f32_restatement(wname):
    """full-size drift localisation: conv stem + positional embedding and the residual stream after blocks 1, 16 and 32 against the
    numpy restatement.  F16 weights: the restatement in its reference-rounding mode.  Q8_0 / Q4_0: plain F32 math on the DEQUANTISED
    weights -- the tight bound (<= 2e-3) that proves block decode + GEMM exact at full size, where the reference itself is only
    comparable to 3e-2 because ggml also quantises the activations"""
    from oracle import encoder_np, mel_np
    buf = mfm.to_bytes(synth.synth_model(synth.FULL_HPARAMS, WT[wname], seed=1234))
    ctx = Context.init_from_buffer(buf)
    orc = encoder_np.EncoderOracle(mfm.read_model(buf), "ggml" if wname == "f16" else "f32")
    del buf
    pcm = synth.synth_pcm(480000, seed=0)
    assert ctx.pcm_to_mel(pcm) == 0
    win = mel_np.window(ctx.get_mel(), 0, 1500)
    taps = {0: None, 1: None, 16: None, 32: None}
    want_emb = orc.encode(win, taps=taps)
    for k in sorted(taps):
        ctx.debug_forward_layers(k)
        assert ctx.encode(0) == 0
        got = ctx.debug_residual(0)
        bound = 1e-3 if k == 0 else 2e-3
        assert rel_l2(got, taps[k]) < bound, (k, rel_l2(got, taps[k]))
    ctx.debug_forward_layers(-1)
    assert ctx.encode(0) == 0
    emb = ctx.get_embeddings()[0]
    t = TOL["f16"] if wname == "f16" else TOL["quant_vs_f32_restatement"]
    assert rel_l2(emb, want_emb) < t["rel_l2"], rel_l2(emb, want_emb)
    assert max_abs(emb, want_emb) < 2 * t["max_abs"]
    ctx.free()


_NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from qwen2_audio_whisper_ggml_b200 import Context, api, ggml_quant as gq, modelfile as mfm, parallel, synth
api.log_set(lambda *_: None)
p = api.default_context_params()
p.gpu_device = int(os.environ["LOCAL_RANK"])
ctx = Context.init_from_buffer(mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=13)), p)
ctx.set_max_batch(4)
win = 2 * synth.TINY_HPARAMS["n_audio_ctx"] * 160
B = 7
pcm = np.stack([synth.synth_pcm(win, seed=300 + w, kind="chirp" if w % 3 else "noise") for w in range(B)])
ns = np.full(B, win, dtype=np.int32); ns[2] = 16000
full, (s, e) = parallel.encode_sharded(lambda w, n: ctx.encode_batch(np.ascontiguousarray(w), n), pcm, ns, gather_to=world - 1)
assert (s, e) == parallel.shard_bounds(B, rank, world)
if rank == world - 1:
    want = ctx.encode_batch(pcm, ns)                      # the whole batch on one device
    assert full.shape == want.shape and np.array_equal(full, want), "NCCL gather order / sharded != single device"
else:
    assert full is None
dist.barrier(); ctx.free(); dist.destroy_process_group(); print("ok", rank)
'''


def test_two_rank_nccl_gather_order_and_sharded_equals_single(tmp_path):
    """one process per GPU (torchrun), windows sharded, NCCL gather to the last rank: ordered by window index and bit-identical to the
    same batch on one device"""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    worker = tmp_path / "nccl_worker.py"
    worker.write_text(_NCCL_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(worker), root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_cli_on_samples_jfk_full_model(tmp_path):
    """BASELINE configs[0] end to end from files, the reference README's `./bin/main -f ./samples/jfk.wav`: the full-size F16 model
    file on disk + the samples/jfk clip (decoded once to 16 kHz WAV, tests/golden/decode_jfk.py) through q2w-main; the 20 values it
    prints are the reference's own print-out for the same file and clip (fullset golden, whisper_print_emb_enc)"""
    import subprocess
    path = os.path.join(GOLD, "fullset_f16.npz")
    exe = os.path.abspath(os.path.join(os.path.dirname(GOLD), "..", "qwen2_audio_whisper_ggml_b200", "q2w-main"))
    if not os.path.exists(path) or not os.path.exists(exe):
        pytest.skip("fullset golden or q2w-main missing")
    model = tmp_path / "qwen2-audio-encoder-f16.bin"
    mfm.save(str(model), synth.synth_model(synth.FULL_HPARAMS, WT["f16"], seed=1234))
    r = subprocess.run([exe, "-m", str(model), "-f", os.path.join(GOLD, "jfk_16k.wav"), "-np"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    got = np.array([float(x) for x in lines[0].split()])
    want = np.load(path)["jfk_first20"]
    assert got.shape == (20,) and np.abs(got - want).max() < 5e-3 + TOL["f16"]["max_abs"] / 2, np.abs(got - want).max()


def test_multi_modal_projector_after_the_path():
    """Linear(n_audio_state -> n_out) + bias on every embedding row (HF Qwen2AudioMultiModalProjector.linear) as one more
    tcgen05 GEMM fed by the F16 rows the pool + final-LayerNorm kernel writes in the same pass; checked against the same Linear in
    F32 on the library's own embeddings (F16-rounded A and W, F32 accumulate), single window, batch, and a width that is not a tile multiple"""
    ctx, _ = tiny_ctx("f16", seed=21)
    D = synth.TINY_HPARAMS["n_audio_state"]
    rng = np.random.default_rng(9)
    pcm = synth.synth_pcm(32000, seed=3)
    assert ctx.full(pcm) == 0
    before = ctx.get_embeddings()
    with pytest.raises(Exception):
        ctx.project()                                               # no projector yet
    for n_out, dt in ((96, np.float32), (264, np.float16)):
        W = (rng.standard_normal((n_out, D)) / np.sqrt(D)).astype(dt)
        b = (0.1 * rng.standard_normal(n_out)).astype(np.float32)
        assert ctx.set_projector(W, b) == 0
        assert ctx.full(pcm) == 0
        emb = ctx.get_embeddings()
        assert np.array_equal(emb, before)                          # the encoder output itself is unchanged by the extra F16 store
        got = ctx.project()
        want = emb[0].astype(np.float16).astype(np.float32) @ W.astype(np.float16).astype(np.float32).T + b
        assert got.shape == (1, emb.shape[1], n_out) and rel_l2(got[0], want) < 2e-5, rel_l2(got[0], want)
    win = 2 * synth.TINY_HPARAMS["n_audio_ctx"] * 160
    batch = np.stack([synth.synth_pcm(win, seed=70 + k, kind="chirp" if k % 2 else "noise") for k in range(5)])
    ctx.set_max_batch(2)
    embs = ctx.encode_batch(batch)
    proj = ctx.project()
    assert proj.shape == (5, embs.shape[1], 264)
    for k in range(5):
        want = embs[k].astype(np.float16).astype(np.float32) @ W.astype(np.float16).astype(np.float32).T + b
        assert rel_l2(proj[k], want) < 2e-5, k
    assert ctx.set_projector(np.zeros((100, D), np.float32)) == -1   # width must be a multiple of 8
    ctx.free()


def test_multi_device_error_paths():
    """a device ordinal that does not exist fails the load (NULL, like any init error); an empty / NULL device list is rejected; freeing a
    multi-device context twice over is safe"""
    buf = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, WT["f16"], seed=13))
    with pytest.raises(Exception):
        Context.init_from_buffer(buf, devices=[0, 63])
    with pytest.raises(Exception):
        Context.init_from_buffer(buf, devices=[])
    p = api.default_context_params()
    p.gpu_device = 63
    with pytest.raises(Exception):
        Context.init_from_buffer(buf, p)
    ctx = Context.init_from_buffer(buf, devices=[0, 0])
    ctx.free()
    ctx.free()
    Context.init_from_buffer(buf).free()                       # the library is still usable afterwards
