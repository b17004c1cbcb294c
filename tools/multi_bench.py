"""One process, all GPUs, through the C API (whisper_encode_batch_multi): end-to-end audio-s/s from pinned host buffers for
B windows per device, with and without the gather onto device 0.  python tools/multi_bench.py [windows_per_gpu] [steps] [wtype]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from qwen2_audio_whisper_ggml_b200 import Context, api, lib as L

lib = L.load_library()
api.log_set(lambda *_: None)
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wtype = sys.argv[3] if len(sys.argv) > 3 else "f16"
p = api.default_context_params()
p.gpu_device = -1
t0 = time.time()
ctx = Context.init_from_buffer(bench.build_model_bytes(wtype), p)
G = ctx.n_devices()
ctx.set_max_batch(per_gpu)
B = per_gpu * G
one = bench.synth_windows(per_gpu, 0)
host = torch.empty((B, 480000), dtype=torch.float32).pin_memory()
for g in range(G):
    host[g * per_gpu:(g + 1) * per_gpu] = torch.from_numpy(one)
out = torch.empty((B, 750, 1280), dtype=torch.float32).pin_memory()
setup = time.time() - t0
res = {"devices": ctx.devices(), "windows_per_gpu": per_gpu, "total_windows": B, "weights": wtype, "setup_s": round(setup, 1)}
for name, gather in (("no_gather", -1), ("gather_to_device_0", 0)):
    for _ in range(2):
        ctx.encode_batch_multi(host.numpy(), out=out.numpy(), gather_device=gather)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.encode_batch_multi(host.numpy(), out=out.numpy(), gather_device=gather)
    dt = (time.perf_counter() - t0) / steps
    mm = api.wlib().whisper_q2w_multi(ctx._h)
    dev_ms = [lib.q2w_multi_last_device_ms(mm, i) for i in range(G)] if mm else []
    res[name] = {"audio_s_per_s": 30.0 * B / dt, "ms_per_call": 1e3 * dt, "per_device_ms_last_call": [round(x, 1) for x in dev_ms]}
out2 = torch.empty_like(out).pin_memory()
outs = [out, out2]
def run_async(n):
    prev = None
    for i in range(n):
        t = ctx.encode_batch_async(host.numpy(), outs[i & 1].numpy())
        if prev is not None:
            ctx.wait(prev)
        prev = t
    ctx.wait(prev)
run_async(2)
t0 = time.perf_counter()
run_async(steps)
dt = (time.perf_counter() - t0) / steps
res["async_two_in_flight"] = {"audio_s_per_s": 30.0 * B / dt, "ms_per_call": 1e3 * dt}
chk = bench.golden_check(out[0].numpy(), wtype)
res["parity_window0"] = {k: chk[k] for k in ("rel_l2", "max_abs", "ok")} if chk.get("checked") else chk
same = all(torch.equal(out[g * per_gpu:(g + 1) * per_gpu], out[:per_gpu]) for g in range(1, G))
res["all_devices_bit_identical"] = bool(same)
if G > 1:
    res["gathered_equals_host"] = bool(np.array_equal(ctx.gathered(B)[::max(1, B // 16)], out.numpy()[::max(1, B // 16)]))
print(json.dumps(res))
ctx.free()
