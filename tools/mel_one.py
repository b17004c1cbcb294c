"""mel kernel alone for an ncu capture: 64 x 30 s windows, the batch path's frame count"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen2_audio_whisper_ggml_b200 import lib as L, synth
import bench
lib = L.load_library()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
filt = synth.slaney_mel_filters(128)
pcm = torch.from_numpy(bench.synth_windows(B, 0)).cuda()
n_frames, ld = 3002, 3008
mel = torch.empty(B, 128, ld, device="cuda")
mx = torch.zeros(B, device="cuda", dtype=torch.int32)
for _ in range(3):
    L.check(lib.q2w_op_mel(filt.ctypes.data, 128, pcm.data_ptr(), 480000, None, 480000, B, n_frames, mel.data_ptr(), ld, mx.data_ptr(), 0, None))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    L.check(lib.q2w_op_mel(filt.ctypes.data, 128, pcm.data_ptr(), 480000, None, 480000, B, n_frames, mel.data_ptr(), ld, mx.data_ptr(), 0, None))
e1.record(); torch.cuda.synchronize()
print(f"mel B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per call (includes plan create + sync in the op shim)")
