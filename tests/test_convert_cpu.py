"""CPU tier: the checkpoint converter, cross-checked against an INDEPENDENT definition of the architecture --
HF transformers' Qwen2AudioEncoder (the audio_tower of Qwen2-Audio) with the tanh GELU ggml uses.  The converted file is
run through the unmodified reference (oracle/_ref) and through the numpy restatement; both must reproduce HF's forward."""
import numpy as np
import pytest

from util import rel_l2

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")

from oracle import encoder_np, mel_np  # noqa: E402
from qwen2_audio_whisper_ggml_b200 import convert, ggml_quant as gq, modelfile as mfm, synth  # noqa: E402


def tiny_hf_encoder(seed=0):
    from transformers.models.qwen2_audio.configuration_qwen2_audio import Qwen2AudioEncoderConfig
    from transformers.models.qwen2_audio.modeling_qwen2_audio import Qwen2AudioEncoder
    cfg = Qwen2AudioEncoderConfig(num_mel_bins=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=512, d_model=128,
                                  max_source_positions=100, activation_function="gelu_pytorch_tanh", dropout=0.0, attention_dropout=0.0,
                                  activation_dropout=0.0, encoder_layerdrop=0.0)
    torch.manual_seed(seed)
    enc = Qwen2AudioEncoder(cfg).eval().float()
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if n.endswith("bias"):
                p.normal_(0.0, 0.02)
            elif "layer_norm" in n and n.endswith("weight"):
                p.copy_(1.0 + 0.02 * torch.randn_like(p))
    return enc


def test_state_dict_keys_are_the_loaders_tensor_names():
    enc = tiny_hf_encoder()
    sd = enc.state_dict()
    hp = convert.hparams_from_state_dict(sd, n_head=2)
    assert set(mfm.expected_shapes(hp)) <= set(sd), set(mfm.expected_shapes(hp)) - set(sd)
    assert (hp["n_audio_ctx"], hp["n_audio_state"], hp["n_audio_layer"], hp["n_mels"]) == (100, 128, 2, 128)


def test_converted_model_reproduces_hf_forward(ref):
    enc = tiny_hf_encoder(1)
    sd = {k: v for k, v in enc.state_dict().items()}
    hp = convert.hparams_from_state_dict(sd, n_head=2, n_vocab=64)
    mf32 = convert.from_state_dict(sd, gq.GGML_TYPE_F32, hp)
    pcm = synth.synth_pcm(32000, seed=4)
    mel = mel_np.log_mel_spectrogram(pcm, mf32.filters)
    win = mel_np.window(mel, 0, 100)
    with torch.no_grad():
        want = enc(torch.from_numpy(win)[None]).last_hidden_state[0].numpy()         # [50, 128]
    # numpy restatement on the F32 file: same math as HF up to summation order
    got = encoder_np.EncoderOracle(mf32, "f32").encode(win)
    assert got.shape == want.shape
    assert rel_l2(got, want) < 1e-4, rel_l2(got, want)      # F32 summation order only
    # the unmodified reference on the converted F32 and F16 files
    for wt, tol in ((gq.GGML_TYPE_F32, 5e-4), (gq.GGML_TYPE_F16, 2e-3)):
        mf = convert.from_state_dict(sd, wt, hp)
        ctx = ref.RefContext(mfm.to_bytes(mf))
        assert ctx.full(pcm) == 0
        r = rel_l2(ctx.get_embeddings(), want)
        assert r < tol, (gq.TYPE_NAMES[wt], r)
        ctx.free()


def test_converter_cli_round_trip(tmp_path):
    enc = tiny_hf_encoder(2)
    ck = tmp_path / "ck.pt"
    torch.save({"dims": dict(n_mels=128, n_audio_ctx=100, n_audio_state=128, n_audio_head=2, n_audio_layer=2, n_vocab=64),
                "model_state_dict": enc.state_dict()}, ck)
    out = tmp_path / "m-q8_0.bin"
    convert.main([str(ck), str(out), "--wtype", "q8_0"])
    mf = mfm.load(str(out))
    assert mf.hparams["ftype"] == 2007 and mf.wtype == gq.GGML_TYPE_Q8_0 and len(mf.tensors) == 7 + 15 * 2
    assert mf.tensor("layers.0.fc1.weight").ttype == gq.GGML_TYPE_Q8_0 and mf.tensor("conv1.weight").ttype == gq.GGML_TYPE_F16
    assert mf.tensor("embed_positions.weight").ttype == gq.GGML_TYPE_F32 and mf.tensor("conv2.bias").ne == (1, 128)
