"""CPU tier (-m "not gpu"): host logic, file format, quantisers pinned against ggml, C-ABI symbol check, struct layout,
world_size-2 gloo test of the window sharding.  No compute call touches the CUDA library here (there is no GPU)."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from qwen2_audio_whisper_ggml_b200 import ggml_quant as gq, modelfile as mfm, parallel, synth
from qwen2_audio_whisper_ggml_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load_library()
    names = L.header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/q2w_b200.h but not exported"
    assert set(L._SIGS) == set(names), set(L._SIGS) ^ set(names)
    assert b"sm_100a" in lib.q2w_build_info()


def test_whisper_api_symbols_exported():
    import re
    from qwen2_audio_whisper_ggml_b200 import api
    lib = api.wlib()
    src = open(os.path.join(ROOT, "include", "qwen2-whisper.h")).read()
    names = set(re.findall(r"WHISPER_API[^;(]*?\b(whisper_\w+)\s*\(", src))
    assert len(names) >= 55
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qwen2-whisper.h but not exported"


def test_no_gpu_fails_loudly_not_silently():
    """without an sm_100 device every compute entry point must return an error, never fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    hp = L.HParams(**synth.TINY_HPARAMS)
    h = ctypes.c_void_p()
    rc = lib.q2w_model_create(ctypes.byref(h), ctypes.byref(hp), 1, 0)
    assert rc == -3 and b"no CPU fallback" in lib.q2w_last_error()
    from qwen2_audio_whisper_ggml_b200 import Context, api
    api.log_set(lambda *_: None)
    with pytest.raises(L.Q2WError):
        Context.init_from_buffer(mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16)))
    api.log_set(None)


def test_default_params_match_reference_values():
    from qwen2_audio_whisper_ggml_b200 import api
    p = api.default_context_params()          # src/qwen2-whisper.cpp:3012-3028
    assert (p.use_gpu, p.flash_attn, p.gpu_device, p.dtw_n_top, p.dtw_mem_size) == (True, False, 0, -1, 128 * 1024 * 1024)
    f = api.wlib().whisper_full_default_params()   # :4231-4295 (the reference forgets to return it)
    assert f.n_max_text_ctx == 16384 and f.no_context and f.print_progress and f.language == b"en"
    assert abs(f.entropy_thold - 2.4) < 1e-6 and abs(f.no_speech_thold - 0.6) < 1e-6 and f.offset_ms == 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_struct_layouts_identical_to_reference_header():
    """sizeof / offsetof of the two by-value structs must match the reference header, or callers break silently"""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include HDR
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(struct whisper_context_params), offsetof(struct whisper_context_params, gpu_device),
         offsetof(struct whisper_context_params, dtw_aheads), offsetof(struct whisper_context_params, dtw_mem_size),
         sizeof(struct whisper_model_loader), offsetof(struct whisper_model_loader, close));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct whisper_full_params), offsetof(struct whisper_full_params, offset_ms),
         offsetof(struct whisper_full_params, thold_pt), offsetof(struct whisper_full_params, audio_ctx),
         offsetof(struct whisper_full_params, suppress_regex), offsetof(struct whisper_full_params, language),
         offsetof(struct whisper_full_params, temperature), offsetof(struct whisper_full_params, abort_callback),
         offsetof(struct whisper_full_params, i_start_rule));
  return 0; }
'''
    outs = []
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        for hdr, inc in (('"qwen2-whisper.h"', [f"-I{ROOT}/include"]),
                         ('"qwen2-whisper.h"', [f"-I{REF}/include", f"-I{REF}/ggml/include"])):
            exe = os.path.join(td, "t")
            subprocess.check_call(["gcc", "-std=c11", f"-DHDR={hdr}", *inc, src, "-o", exe])
            outs.append(subprocess.check_output([exe]).decode())
    assert outs[0] == outs[1], outs
    from qwen2_audio_whisper_ggml_b200 import api
    assert int(outs[0].split()[0]) == ctypes.sizeof(api.ContextParams)
    assert int(outs[0].splitlines()[1].split()[0]) == ctypes.sizeof(api.FullParams)


def test_quantisers_bit_exact_vs_ggml(ref):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((48, 1280)) * 0.05).astype(np.float32)
    x[3, :32] = 0.0                      # all-zero block: d = 0, id = 0
    x[5, 40] = 7.0                       # outlier
    x[6, 64:96] = -x[6, 64:96].max()     # tie on |max| with a negative first occurrence (q4_0 sign rule)
    x[7, 96:128] = np.float32(0.5)       # roundf half-away cases after scaling
    for t in (gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0):
        mine = gq.quantize(x, t).reshape(-1)
        theirs = ref.ref_quantize(x, t)
        assert np.array_equal(mine, theirs), gq.TYPE_NAMES[t]
        assert np.array_equal(gq.dequantize(mine, t, 1280).reshape(-1), ref.ref_dequantize(theirs, t, x.size))


def test_quantisers_bit_exact_vs_committed_ggml_blocks():
    """the same pin without oracle/_ref: blocks produced once by the reference's ggml_quantize_chunk / dequantize_row_* (inputs with an
    all-zero block, an outlier, a negative |max| tie and half-way roundings), committed as tests/golden/ggml_quant_blocks.npz"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ggml_quant_blocks.npz"))
    x = g["x"]
    for t in (gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0):
        nm = gq.TYPE_NAMES[t]
        mine = gq.quantize(x, t).reshape(-1)
        assert np.array_equal(mine, g[f"raw_{nm}"]), nm
        assert np.array_equal(gq.dequantize(mine, t, x.shape[1]).reshape(-1), g[f"deq_{nm}"]), nm


def test_model_file_round_trip_and_reference_loads_it(ref):
    for wt in (gq.GGML_TYPE_F16, gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0, gq.GGML_TYPE_F32):
        mf = synth.synth_model(synth.TINY_HPARAMS, wt, seed=5)
        buf = mfm.to_bytes(mf)
        back = mfm.read_model(buf)
        assert back.hparams == mf.hparams and len(back.tensors) == 7 + 15 * 2
        assert all(np.array_equal(a.data, b.data) and a.ne == b.ne and a.ttype == b.ttype for a, b in zip(mf.tensors, back.tensors))
        assert mfm.to_bytes(back) == buf
        ctx = ref.RefContext(buf)            # the unmodified reference loader accepts the file
        ctx.free()
    # F16 file -> quantised file equals building the quantised file directly (the fork's converter -> quantiser pipeline)
    f16 = synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=5)
    for wt in (gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0):
        assert mfm.to_bytes(mfm.quantize_model(f16, wt)) == mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, wt, seed=5))


def test_full_size_file_sizes_match_survey():
    """byte counts of the three full-size files (SURVEY 8c probe) from the format arithmetic alone"""
    hp = synth.FULL_HPARAMS
    sizes = {}
    for wt in (gq.GGML_TYPE_F16, gq.GGML_TYPE_Q8_0, gq.GGML_TYPE_Q4_0):
        n = 4 + 44 + 8 + 4 * 128 * 201 + 4
        for name, ne in mfm.expected_shapes(hp).items():
            tt = mfm.tensor_type_for(name, ne, wt)
            n += 12 + 4 * len(ne) + len(name) + gq.row_bytes(tt, ne[0]) * int(np.prod(ne[1:]))
        sizes[wt] = n
    assert sizes[gq.GGML_TYPE_F16] == 1278896820 and sizes[gq.GGML_TYPE_Q8_0] == 689072820 and sizes[gq.GGML_TYPE_Q4_0] == 374500020


def test_mel_filterbank_matches_whisper():
    f = synth.slaney_mel_filters(128)
    assert f.shape == (128, 201) and int((f != 0).sum()) == 394 and int((f != 0).sum(axis=1).max()) <= 9
    try:
        from transformers import WhisperFeatureExtractor
    except Exception:
        pytest.skip("transformers not importable")
    ref = WhisperFeatureExtractor(feature_size=128).mel_filters.T
    assert np.abs(f - ref).max() < 1e-7


def test_shard_bounds_partition():
    for B in (0, 1, 5, 64, 120, 256, 257):
        for G in (1, 2, 3, 4, 8):
            bounds = [parallel.shard_bounds(B, r, G) for r in range(G)]
            assert bounds[0][0] == 0 and bounds[-1][1] == B
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(G - 1))
            sizes = [e - s for s, e in bounds]
            assert max(sizes) - min(sizes) <= 1
            for r, (s, e) in enumerate(bounds):
                assert all(w * G // B == r for w in range(s, e))      # window w -> rank floor(w * G / B)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from qwen2_audio_whisper_ggml_b200 import parallel
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
B = 7
windows = np.arange(B * 4, dtype=np.float32).reshape(B, 4)
ns = np.arange(B, dtype=np.int32) + 1
def fake_encode(w, n):      # stands in for Context.encode_batch: result depends only on the window's own data
    return (w.sum(axis=1)[:, None, None] + n[:, None, None] * np.ones((1, 3, 2), np.float32)).astype(np.float32)
local, (s, e) = parallel.encode_sharded(fake_encode, windows, ns)
assert (s, e) == parallel.shard_bounds(B, dist.get_rank(), 2) and local.shape == (e - s, 3, 2)
full, _ = parallel.encode_sharded(fake_encode, windows, ns, gather_to=1)
if dist.get_rank() == 1:
    assert np.array_equal(full, fake_encode(windows, ns)), "gather order"
else:
    assert full is None
dist.barrier(); dist.destroy_process_group(); print("ok")
'''


def test_sharding_world_size_2_gloo():
    port = 29500 + os.getpid() % 2000
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_GLOO_WORKER)
    try:
        procs = [subprocess.Popen([sys.executable, f.name, ROOT, str(port), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                 for r in range(2)]
        outs = [p.communicate(timeout=240)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
    finally:
        os.unlink(f.name)
