"""one GEMM shape for ncu: python tools/gemm_one.py M N K epi"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L
lib = L.load_library()
m, n, k, epi = [int(x) for x in sys.argv[1:5]]
g = torch.Generator(device="cuda").manual_seed(0)
A = (torch.randn(m, k, device="cuda", generator=g) * 0.5).half()
W = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
bias = torch.randn(n, device="cuda", generator=g)
out = torch.zeros(m, n, device="cuda", dtype=torch.half if epi in (0, 1) else torch.float32)
pos = torch.randn(1500, n, device="cuda", generator=g)
for _ in range(4):
    L.check(lib.q2w_op_gemm(A.data_ptr(), k, W.data_ptr(), k, m, n, k, bias.data_ptr(), out.data_ptr(), n, epi,
                            out.data_ptr() if epi == 2 else None, pos.data_ptr(), 1500, n // 2, 0.125, None))
torch.cuda.synchronize()
print("ok")
