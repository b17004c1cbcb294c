"""a few single-window forwards (eager, no graph because profiling hooks are on) for an ncu launch list"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from qwen2_audio_whisper_ggml_b200 import Context, api, lib as L
lib = L.load_library(); api.log_set(lambda *_: None)
ctx = Context.init_from_buffer(bench.build_model_bytes("f16")); ctx.set_max_batch(1)
L.check(lib.q2w_profile_enable(ctx.q2w_state(), 1))      # forces the eager path so every kernel is a separate launch
dev = torch.from_numpy(bench.synth_windows(1, 0)).cuda(); torch.cuda.synchronize()
for _ in range(3):
    ctx.encode_batch_device(dev.data_ptr(), 480000, 1)
torch.cuda.synchronize(); print("ok")
