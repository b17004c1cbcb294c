"""one launch of the attention kernel at the bench shape (diagnostic timeline build prints from the kernel)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L
lib = L.load_library()
B, H, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 20, 1500
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.5).half()
o = torch.empty(B * T, H * 64, device="cuda", dtype=torch.half)
L.check(lib.q2w_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
torch.cuda.synchronize()
