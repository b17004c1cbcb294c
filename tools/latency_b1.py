"""B = 1 latency breakdown by kernel class (profile hooks; eager launches), plus the graph-replay p50"""
import sys, os, ctypes as C, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from qwen2_audio_whisper_ggml_b200 import Context, api, lib as L
lib = L.load_library()
api.log_set(lambda *_: None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = Context.init_from_buffer(bench.build_model_bytes("f16"))
ctx.set_max_batch(max(B, 1))
st = ctx.q2w_state()
dev = torch.from_numpy(bench.synth_windows(B, 0)).cuda()
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(lib.q2w_state_stream(st))
for _ in range(4):
    ctx.encode_batch_device(dev.data_ptr(), 480000, B)
lat = []
for _ in range(30):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream); ctx.encode_batch_device(dev.data_ptr(), 480000, B); a1.record(stream); torch.cuda.synchronize()
    lat.append(a0.elapsed_time(a1))
print(f"B={B} p50 {statistics.median(lat):.3f} ms  min {min(lat):.3f}")
L.check(lib.q2w_profile_enable(st, 1))
reps = 10
for _ in range(reps):
    ctx.encode_batch_device(dev.data_ptr(), 480000, B)
tot = 0
for ci, nm in enumerate(["gemm", "attention", "layernorm", "mel", "im2col", "dequant"]):
    ms, cnt, fl, by = C.c_double(), C.c_long(), C.c_double(), C.c_double()
    L.check(lib.q2w_profile_read(st, ci, C.byref(ms), C.byref(cnt), C.byref(fl), C.byref(by)))
    if cnt.value:
        print(f"  {nm:10s} {ms.value / reps:7.3f} ms/pass  {cnt.value // reps:4d} launches  avg {1e3 * ms.value / cnt.value:6.1f} us" + (f"  {fl.value / ms.value / 1e9:7.1f} TFLOP/s" if fl.value else ""))
        tot += ms.value / reps
print(f"  sum of kernel classes {tot:.3f} ms")
