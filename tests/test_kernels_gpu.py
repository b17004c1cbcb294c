"""-m gpu: every sm_100a kernel against a plain torch FP32 restatement of the same op, through the C-ABI shims."""
import ctypes as C

import numpy as np
import pytest

from util import max_abs, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from qwen2_audio_whisper_ggml_b200 import ggml_quant as gq  # noqa: E402
from qwen2_audio_whisper_ggml_b200 import lib as L  # noqa: E402
from qwen2_audio_whisper_ggml_b200 import synth  # noqa: E402


def ck(rc):
    L.check(rc)
    torch.cuda.synchronize()


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


GEMM_SHAPES = [(300, 384, 128), (128, 256, 64), (1500, 1280, 1280), (4113, 3840, 1280), (1500, 1280, 5120), (3000, 1280, 384),
               (777, 128, 512), (200, 136, 240)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_gemm(lib, M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    acc = A.float() @ W.float().t() + bias
    period = 100
    pos = torch.randn(period, N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    scale_cols, scale = (N // 2) // 8 * 8, 0.125
    if epi == L.EPI_BIAS_F16:
        want = acc.clone()
        want[:, :scale_cols] *= scale
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
    elif epi == L.EPI_BIAS_GELU_F16:
        want = gelu_tanh(acc)
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
    elif epi == L.EPI_BIAS_RESID_F32:
        want = acc + resid
        out = resid.clone()          # in place, as the engine uses it
    elif epi == L.EPI_BIAS_GELU_POS_F32:
        want = gelu_tanh(acc) + pos[torch.arange(M, device="cuda") % period]
        out = torch.full((M, N), float("nan"), device="cuda")
    else:
        want = acc
        out = torch.full((M, N), float("nan"), device="cuda")
    ck(lib.q2w_op_gemm(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), out.data_ptr(), N, epi,
                       out.data_ptr() if epi == L.EPI_BIAS_RESID_F32 else None, pos.data_ptr(), period, scale_cols, scale, None))
    got = out.float().cpu().numpy()
    want = want.cpu().numpy()
    assert np.isfinite(got).all()
    tol = 2e-3 if out.dtype == torch.half else 2e-5
    assert rel_l2(got, want) < tol, (rel_l2(got, want), max_abs(got, want))


@pytest.fixture
def single_pass_gemm(lib):
    """pin the bit-reproducible single-pass residual epilogue (no split-K) for tests that compare outputs bit for bit"""
    lib.q2w_op_set_gemm_splitk(0)
    yield
    lib.q2w_op_set_gemm_splitk(-1)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K", [(1500, 1280, 1280), (1500, 1280, 5120), (700, 1280, 5120), (257, 256, 2048), (100, 1280, 5120), (1500, 1288, 1096)])
def test_gemm_residual_split_k(lib, mode, M, N, K):
    """small-M residual epilogue with the K range cut over several CTA pairs (each partial is a TMA reduction store into x, the
    bias rides on the segment that holds k-block 0): same answer as the single pass up to the order of the F32 adds, on top of a
    NaN-free residual; rows >= M / columns >= N of a wider buffer stay untouched"""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + mode)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = torch.randn(M + 3, N + 8, device="cuda", generator=g)
    want = resid.clone()
    want[:M, :N] += A.float() @ W.float().t() + bias
    outs = []
    try:
        for md in (0, mode, mode):
            lib.q2w_op_set_gemm_splitk(md)
            out = resid.clone()
            ck(lib.q2w_op_gemm(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), out.data_ptr(), N + 8, L.EPI_BIAS_RESID_F32,
                               out.data_ptr(), None, 0, 0, 1.0, None))
            outs.append(out)
    finally:
        lib.q2w_op_set_gemm_splitk(-1)
    for out in outs:
        assert torch.equal(out[M:], resid[M:]) and torch.equal(out[:, N:], resid[:, N:])
        assert rel_l2(out.cpu().numpy(), want.cpu().numpy()) < 2e-5
    # split and single pass agree to F32 rounding of a handful of adds (not bit for bit: the L2 adds arrive in any order)
    assert max_abs(outs[1].cpu().numpy(), outs[0].cpu().numpy()) < 2e-5
    assert max_abs(outs[2].cpu().numpy(), outs[1].cpu().numpy()) < 2e-5


@pytest.mark.parametrize("ttype", [gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0])
@pytest.mark.parametrize("M,N,K,epi", [(300, 384, 128, 0), (1500, 1280, 1280, 2), (2100, 5120, 1280, 1), (777, 1280, 5120, 2), (520, 136, 192, 4)])
def test_gemm_in_kernel_dequant(lib, single_pass_gemm, ttype, M, N, K, epi):
    """W stays in ggml blocks; the GEMM's decode warpgroup must reproduce dequantize_row_* + F16 rounding exactly, so the result
    equals the F16-weight GEMM on the decoded matrix bit for bit"""
    rng = np.random.default_rng(M + N + K + ttype)
    w = (rng.standard_normal((N, K)) / K ** 0.5).astype(np.float32)
    w[1, :32] = 0.0
    raw = gq.quantize(w, ttype).reshape(-1)
    wdq = torch.from_numpy(gq.dequantize(raw, ttype, K).astype(np.float16)).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    d_raw = torch.from_numpy(raw.copy()).cuda()
    dt = torch.half if epi in (0, 1) else torch.float32
    resid = torch.randn(M, N, device="cuda", generator=g)
    out_q = resid.clone().to(dt) if epi == 2 else torch.zeros(M, N, device="cuda", dtype=dt)
    out_f = out_q.clone()
    ck(lib.q2w_op_gemm_q(A.data_ptr(), K, d_raw.data_ptr(), ttype, M, N, K, bias.data_ptr(), out_q.data_ptr(), N, epi,
                         out_q.data_ptr() if epi == 2 else None, N // 2 // 8 * 8, 0.125, None))
    ck(lib.q2w_op_gemm(A.data_ptr(), K, wdq.data_ptr(), K, M, N, K, bias.data_ptr(), out_f.data_ptr(), N, epi,
                       out_f.data_ptr() if epi == 2 else None, None, 0, N // 2 // 8 * 8, 0.125, None))
    assert torch.equal(out_q, out_f)
    want = A.float() @ wdq.float().t() + bias
    if epi == 4:
        assert rel_l2(out_q.cpu().numpy(), want.cpu().numpy()) < 2e-5


def test_gemm_strided_operands(lib):
    """lda / ldw / ldo larger than the logical widths (views into wider buffers)"""
    M, N, K = 520, 256, 192
    g = torch.Generator(device="cuda").manual_seed(5)
    Abig = (torch.randn(M, K + 64, device="cuda", generator=g)).half()
    Wbig = (torch.randn(N, K + 8, device="cuda", generator=g) / K ** 0.5).half()
    out = torch.zeros(M, N + 16, device="cuda")
    ck(lib.q2w_op_gemm(Abig.data_ptr(), K + 64, Wbig.data_ptr(), K + 8, M, N, K, None, out.data_ptr(), N + 16, L.EPI_BIAS_F32,
                       None, None, 0, 0, 1.0, None))
    want = Abig[:, :K].float() @ Wbig[:, :K].float().t()
    assert rel_l2(out[:, :N].cpu().numpy(), want.cpu().numpy()) < 2e-5
    assert float(out[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,D", [(1, 128), (37, 384), (1500, 1280), (4099, 1280)])
def test_layernorm(lib, M, D):
    g = torch.Generator(device="cuda").manual_seed(M + D)
    x = torch.randn(M, D, device="cuda", generator=g) * 3 + 1.5
    gam = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    bet = 0.1 * torch.randn(D, device="cuda", generator=g)
    y = torch.empty(M, D, device="cuda", dtype=torch.half)
    ck(lib.q2w_op_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), y.data_ptr(), M, D, 1e-5, None))
    want = torch.nn.functional.layer_norm(x.double(), (D,), gam.double(), bet.double(), 1e-5)
    assert max_abs(y.float().cpu().numpy(), want.cpu().numpy()) < 4e-3     # f16 output rounding at |y| < 8
    assert rel_l2(y.float().cpu().numpy(), want.cpu().numpy()) < 5e-4


@pytest.mark.parametrize("B,T,D", [(1, 100, 128), (3, 1500, 1280)])
def test_pool_layernorm(lib, B, T, D):
    g = torch.Generator(device="cuda").manual_seed(B + T)
    x = torch.randn(B * T, D, device="cuda", generator=g) * 2
    gam = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    bet = 0.1 * torch.randn(D, device="cuda", generator=g)
    y = torch.empty(B * T // 2, D, device="cuda")
    ck(lib.q2w_op_pool_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), y.data_ptr(), B, T, D, 1e-5, None))
    xp = (x.view(B, T // 2, 2, D).double().sum(2) / 2).view(-1, D)
    want = torch.nn.functional.layer_norm(xp, (D,), gam.double(), bet.double(), 1e-5)
    assert max_abs(y.cpu().numpy(), want.cpu().numpy()) < 2e-5


# the last four shapes give every persistent CTA several work items (1-, 2-, 3- and 12-tile KV sequences)
@pytest.mark.parametrize("B,T,H", [(1, 64, 2), (2, 100, 2), (1, 1500, 20), (3, 333, 4), (40, 100, 20), (12, 129, 20), (16, 300, 20), (8, 1500, 20)])
def test_attention(lib, B, T, H):
    D = 64 * H
    g = torch.Generator(device="cuda").manual_seed(B * 100 + T + H)
    qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.35                      # pre-scaled queries (the QKV epilogue folds 1/sqrt(64))
    qkv = qkv.half()
    out = torch.full((B * T, D), float("nan"), device="cuda", dtype=torch.half)
    ck(lib.q2w_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, None))
    f = qkv.float().view(B, T, 3, H, 64)
    q, k, v = f[:, :, 0].permute(0, 2, 1, 3), f[:, :, 1].permute(0, 2, 1, 3), f[:, :, 2].permute(0, 2, 1, 3)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    want = (p @ v).permute(0, 2, 1, 3).reshape(B * T, D)
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_l2(got, want.cpu().numpy()) < 2e-3, rel_l2(got, want.cpu().numpy())


def test_attention_growing_scores_take_the_rescale_path(lib):
    """keys whose scores grow along the sequence: the running row max rises by far more than 2^8 between KV tiles, so the lazy
    O rescale (tcgen05.ld -> mul -> tcgen05.st) runs on most tiles, across several items per CTA"""
    B, T, H = 6, 700, 20
    D = 64 * H
    g = torch.Generator(device="cuda").manual_seed(77)
    f = torch.randn(B, T, 3, H, 64, device="cuda", generator=g)
    f[:, :, 0] = f[:, :, 0].abs() * 0.5                                    # positive queries ...
    ramp = torch.linspace(0.05, 3.0, T, device="cuda").view(1, T, 1, 1)
    f[:, :, 1] = f[:, :, 1].abs() * ramp                                    # ... against keys that grow with position: scores up to ~100
    qkv = f.reshape(B * T, 3 * D).half()
    out = torch.full((B * T, D), float("nan"), device="cuda", dtype=torch.half)
    ck(lib.q2w_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, None))
    ff = qkv.float().view(B, T, 3, H, 64)
    q, k, v = ff[:, :, 0].permute(0, 2, 1, 3), ff[:, :, 1].permute(0, 2, 1, 3), ff[:, :, 2].permute(0, 2, 1, 3)
    s_ = q @ k.transpose(-1, -2)
    assert float((s_.max(-1).values - s_[..., :128].max(-1).values).median()) > 20.0   # the max really moves between tiles
    want = (torch.softmax(s_, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * T, D)
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_l2(got, want.cpu().numpy()) < 2e-3, rel_l2(got, want.cpu().numpy())


@pytest.mark.parametrize("ttype", [gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0, gq.GGML_TYPE_F32])
def test_dequant_bit_exact(lib, ttype):
    rng = np.random.default_rng(ttype)
    rows, K = 96, 1280
    w = (rng.standard_normal((rows, K)) * 0.05).astype(np.float32)
    w[0, :32] = 0.0
    raw = gq.quantize(w, ttype).reshape(-1)
    want = gq.dequantize(raw, ttype, K).astype(np.float16)        # decode in F32, one rounding to F16
    src = torch.from_numpy(raw.copy()).cuda()
    dst = torch.empty(rows, K, device="cuda", dtype=torch.half)
    ck(lib.q2w_op_dequant(src.data_ptr(), ttype, dst.data_ptr(), rows, K, None))
    assert np.array_equal(dst.cpu().numpy().view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("ttype", [gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0])
def test_gemm_in_kernel_dequant_split_k(lib, ttype):
    """the decode warpgroup follows the same (tile, k-range) segments as the producer when the K range is cut"""
    M, N, K = 1500, 1280, 5120
    rng = np.random.default_rng(ttype)
    w = (rng.standard_normal((N, K)) / K ** 0.5).astype(np.float32)
    raw = gq.quantize(w, ttype).reshape(-1)
    wdq = torch.from_numpy(gq.dequantize(raw, ttype, K).astype(np.float16)).cuda()
    g = torch.Generator(device="cuda").manual_seed(4)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = torch.randn(M, N, device="cuda", generator=g)
    d_raw = torch.from_numpy(raw.copy()).cuda()
    out = resid.clone()
    ck(lib.q2w_op_gemm_q(A.data_ptr(), K, d_raw.data_ptr(), ttype, M, N, K, bias.data_ptr(), out.data_ptr(), N, 2, out.data_ptr(), 0, 1.0, None))
    want = resid + A.float() @ wdq.float().t() + bias
    assert rel_l2(out.cpu().numpy(), want.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("offset,n_valid,normalise", [(0, 6000, 1), (2900, 3100, 1), (17, 230, 0), (5990, 6000, 1), (6000, 6000, 0)])
def test_conv1_operand(lib, offset, n_valid, normalise):
    """window slice [offset, offset + 2 n_ctx) + zero fill past n_len (src/qwen2-whisper.cpp:2274-2283) + max(x, max - 8), (x + 4) / 4
    (:2643-2649, fused here) + im2col k3 s1 p1 in ggml column order ic * 3 + k, for two windows with different maxima"""
    B, n_mel, n_ctx2, ld = 2, 128, 3000, 6016
    g = torch.Generator(device="cuda").manual_seed(offset + n_valid)
    mel = torch.randn(B, n_mel, ld, device="cuda", generator=g) * 2.0 - 3.0
    mx = mel[:, :, :n_valid].amax(dim=(1, 2))
    bits = mx.view(torch.int32)
    keys = torch.where(bits >= 0, bits, bits ^ 0x7FFFFFFF).contiguous()                 # order-preserving int key (mel.cu float_to_key)
    A1 = torch.full((B * n_ctx2, 3 * n_mel), float("nan"), device="cuda", dtype=torch.half)
    ck(lib.q2w_op_conv1_operand(mel.data_ptr(), ld, n_valid, n_mel, keys.data_ptr(), normalise, offset, n_ctx2, B, A1.data_ptr(), None))
    win = torch.zeros(B, n_mel, n_ctx2 + 2, device="cuda")                               # one zero column of conv padding either side
    n_take = max(0, min(n_ctx2, n_valid - offset))
    src = mel[:, :, offset:offset + n_take]
    if normalise:
        src = (torch.maximum(src, (mx - 8.0).view(B, 1, 1)) + 4.0) / 4.0
    win[:, :, 1:1 + n_take] = src
    cols = torch.stack([win[:, :, k:k + n_ctx2] for k in range(3)], dim=-1)              # [B, n_mel, n_ctx2, 3]
    want = cols.permute(0, 2, 1, 3).reshape(B * n_ctx2, 3 * n_mel).half()
    assert torch.equal(A1, want)


@pytest.mark.parametrize("ttype", [gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0])
@pytest.mark.parametrize("with_attention,B,T,H", [(0, 0, 0, 0), (1, 1, 1500, 20), (1, 3, 333, 4), (1, 1, 64, 2)])
def test_dequant_multi_and_attention_rider(lib, ttype, with_attention, B, T, H):
    """four matrices of different sizes decoded in one go -- by the stand-alone kernel, and by the idle warps of an attention launch (grid
    of 240 / 36 / 1 CTAs): bit-exact against dequantize_row_* + one F16 rounding, and the attention result itself is untouched"""
    rng = np.random.default_rng(ttype + T)
    shapes = [(384, 128), (128, 128), (512, 128), (130, 512)]
    raws, wants, d_src, d_dst = [], [], [], []
    for (n, k) in shapes:
        w = (rng.standard_normal((n, k)) * 0.05).astype(np.float32)
        raw = gq.quantize(w, ttype).reshape(-1)
        wants.append(gq.dequantize(raw, ttype, k).astype(np.float16).reshape(n, k))
        d_src.append(torch.from_numpy(raw.copy()).cuda())
        d_dst.append(torch.full((n, k), float("nan"), device="cuda", dtype=torch.half))
    PtrArr = C.c_void_p * 4
    src = PtrArr(*[t.data_ptr() for t in d_src])
    dst = PtrArr(*[t.data_ptr() for t in d_dst])
    nb = (C.c_ulonglong * 4)(*[n * k // 32 for (n, k) in shapes])
    qkv = out = ref = None
    if with_attention:
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = (torch.randn(B * T, 3 * 64 * H, device="cuda", generator=g) * 0.5).half()
        out = torch.full((B * T, 64 * H), float("nan"), device="cuda", dtype=torch.half)
        ref = torch.empty_like(out)
        ck(lib.q2w_op_attention(qkv.data_ptr(), ref.data_ptr(), B, T, H, None))
    ck(lib.q2w_op_dequant_multi(src, dst, nb, 4, ttype, with_attention, qkv.data_ptr() if with_attention else None,
                                out.data_ptr() if with_attention else None, B, T, H, None))
    for got, want in zip(d_dst, wants):
        assert np.array_equal(got.cpu().numpy().view(np.uint16), want.view(np.uint16))
    if with_attention:
        assert torch.equal(out, ref)


def test_conv2_im2col(lib):
    B, T2, Cc = 2, 200, 128
    g = torch.Generator(device="cuda").manual_seed(9)
    h1 = torch.randn(B * T2, Cc, device="cuda", generator=g).half()
    A2 = torch.empty(B * T2 // 2, 3 * Cc, device="cuda", dtype=torch.half)
    ck(lib.q2w_op_conv2_im2col(h1.data_ptr(), A2.data_ptr(), B, T2, Cc, None))
    x = h1.view(B, T2, Cc).permute(0, 2, 1).float()                                   # [B, C, T2]
    cols = torch.nn.functional.unfold(x.unsqueeze(2), (1, 3), padding=(0, 1), stride=(1, 2))   # [B, C*3, T]
    want = cols.permute(0, 2, 1).reshape(B * T2 // 2, 3 * Cc)
    assert torch.equal(A2.float(), want)


@pytest.mark.parametrize("kind,n", [("chirp", 480000), ("tones", 480000), ("silence", 16000), ("noise", 123457), ("chirp", 201)])
def test_mel_vs_numpy_oracle(lib, kind, n):
    from oracle import mel_np
    filt = synth.slaney_mel_filters(128)
    pcm = synth.synth_pcm(n, seed=11, kind=kind)
    n_len, _ = mel_np.mel_dims(n)
    ld = (n_len + 3) // 4 * 4
    d_pcm = torch.from_numpy(pcm).cuda()
    d_mel = torch.full((128, ld), float("nan"), device="cuda")
    d_max = torch.zeros(1, device="cuda", dtype=torch.int32)
    ck(lib.q2w_op_mel(filt.ctypes.data, 128, d_pcm.data_ptr(), 0, None, n, 1, n_len, d_mel.data_ptr(), ld, d_max.data_ptr(), 1, None))
    got = d_mel[:, :n_len].cpu().numpy()
    want = mel_np.log_mel_spectrogram(pcm, filt)
    assert got.shape == want.shape
    assert max_abs(got, want) < 2e-4, max_abs(got, want)
    assert rel_l2(got, want) < 1e-5


def test_mel_dense_random_filterbank(lib):
    """the filter matrix comes from the model file: any dense matrix must be applied exactly, not just banded triangles"""
    from oracle import mel_np
    rng = np.random.default_rng(3)
    filt = (rng.random((128, 201)) * 0.01).astype(np.float32)
    filt[5] = 0.0
    pcm = synth.synth_pcm(48000, seed=2, kind="noise")
    n_len, _ = mel_np.mel_dims(pcm.size)
    ld = (n_len + 3) // 4 * 4
    d_pcm = torch.from_numpy(pcm).cuda()
    d_mel = torch.empty(128, ld, device="cuda")
    d_max = torch.zeros(1, device="cuda", dtype=torch.int32)
    ck(lib.q2w_op_mel(filt.ctypes.data, 128, d_pcm.data_ptr(), 0, None, pcm.size, 1, n_len, d_mel.data_ptr(), ld, d_max.data_ptr(), 1, None))
    want = mel_np.log_mel_spectrogram(pcm, filt)
    assert max_abs(d_mel[:, :n_len].cpu().numpy(), want) < 2e-4
