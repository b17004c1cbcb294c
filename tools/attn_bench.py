"""attention-only driver for ncu: a few launches of the tcgen05 attention at the bench shape (B windows)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L
lib = L.load_library()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, T = 20, 1500
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.5).half()
o = torch.empty(B * T, H * 64, device="cuda", dtype=torch.half)
for _ in range(3):
    L.check(lib.q2w_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    L.check(lib.q2w_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"attention B={B}: {ms:.3f} ms  {4.0 * B * T * T * H * 64 / ms / 1e9:.1f} TFLOP/s")
