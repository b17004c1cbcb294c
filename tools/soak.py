"""Race hunt: repeat the same calls many times and require bit-identical results where the path is deterministic.
  1. Q8_0 full-size model, single window, split-K pinned off: the rider decode inside the attention kernel feeds the next GEMMs -- a GEMM
     that started before its weights were decoded would change bits.  Graph replay and eager.
  2. F16 tiny model, three replicas on device 0 driven by the multi-device worker threads, gather on: 200 calls.
  3. F16 full-size, 8-window batches through the asynchronous host API with two batches in flight: 40 calls.
python tools/soak.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["Q2W_GEMM_SPLITK"] = "0"
import numpy as np, torch
import bench
from qwen2_audio_whisper_ggml_b200 import Context, api, ggml_quant as gq, lib as L, modelfile as mfm, synth

api.log_set(lambda *_: None)
lib = L.load_library()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300

ctx = Context.init_from_buffer(bench.build_model_bytes("q8_0"))
ctx.set_max_batch(1)
pcm = synth.synth_pcm(480000, seed=0)
assert ctx.pcm_to_mel(pcm) == 0
ref = None
bad = 0
for i in range(N):
    assert ctx.encode(0) == 0
    e = ctx.get_embeddings()
    if ref is None:
        ref = e.copy()
    elif not np.array_equal(e, ref):
        bad += 1
print(f"1. q8_0 single window x {N} (graph replay): {bad} mismatches", flush=True)
L.check(lib.q2w_profile_enable(ctx.q2w_state(), 1))      # eager launches
bad_e = sum(0 if (ctx.encode(0) == 0 and np.array_equal(ctx.get_embeddings(), ref)) else 1 for _ in range(max(10, N // 10)))
print(f"   eager x {max(10, N // 10)}: {bad_e} mismatches", flush=True)
ctx.free()

buf = mfm.to_bytes(synth.synth_model(synth.TINY_HPARAMS, gq.GGML_TYPE_F16, seed=13))
multi = Context.init_from_buffer(buf, devices=[0, 0, 0])
multi.set_max_batch(4)
win = 2 * synth.TINY_HPARAMS["n_audio_ctx"] * 160
batch = np.stack([synth.synth_pcm(win, seed=300 + w, kind="chirp" if w % 3 else "noise") for w in range(11)])
want = multi.encode_batch_multi(batch, gather_device=0).copy()
bad_m = 0
for i in range(200):
    out = multi.encode_batch_multi(batch, gather_device=0)
    if not (np.array_equal(out, want) and np.array_equal(multi.gathered(11), want)):
        bad_m += 1
print(f"2. three replicas / worker threads / gather x 200: {bad_m} mismatches", flush=True)
multi.free()

ctx = Context.init_from_buffer(bench.build_model_bytes("f16"))
ctx.set_max_batch(8)
host = torch.from_numpy(bench.synth_windows(8, 0)).pin_memory()
outs = [torch.empty((8, 750, 1280), dtype=torch.float32).pin_memory() for _ in range(2)]
want = ctx.encode_batch(host.numpy()).copy()
bad_a = 0
prev = None
for i in range(40):
    t = ctx.encode_batch_async(host.numpy(), outs[i & 1].numpy())
    if prev is not None:
        ctx.wait(prev[0])
        if not np.array_equal(outs[prev[1]].numpy(), want):
            bad_a += 1
    prev = (t, i & 1)
ctx.wait(prev[0])
bad_a += 0 if np.array_equal(outs[prev[1]].numpy(), want) else 1
print(f"3. f16 8-window async batches x 40: {bad_a} mismatches", flush=True)
ctx.free()
sys.exit(1 if (bad or bad_e or bad_m or bad_a) else 0)
