#!/bin/bash
# single-window latency with quantised weights: decode-to-scratch + F16 GEMM vs decode inside the GEMM
for wt in q8_0 q4_0; do for f in 0 1; do
python - <<PY
import os, sys, statistics
os.environ["Q2W_FUSED_DEQUANT"] = "$f"
sys.path.insert(0, ".")
import torch, bench
from qwen2_audio_whisper_ggml_b200 import Context, api, lib as L
lib = L.load_library(); api.log_set(lambda *_: None)
ctx = Context.init_from_buffer(bench.build_model_bytes("$wt")); ctx.set_max_batch(1)
dev = torch.from_numpy(bench.synth_windows(1, 0)).cuda(); torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(lib.q2w_state_stream(ctx.q2w_state()))
lat = []
for i in range(33):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream); ctx.encode_batch_device(dev.data_ptr(), 480000, 1); a1.record(stream); torch.cuda.synchronize()
    if i >= 3: lat.append(a0.elapsed_time(a1))
print("$wt fused=$f  B=1 p50 %.3f ms  min %.3f" % (statistics.median(lat), min(lat)), flush=True)
PY
done; done
