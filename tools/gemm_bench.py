"""Per-shape timing of the tcgen05 GEMM (and attention) at the bench problem size, CUDA events on the default stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L

lib = L.load_library()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = B * 1500
shapes = [("qkv   epi0", M, 3840, 1280, 0), ("out   epi2", M, 1280, 1280, 2), ("fc1   epi1", M, 5120, 1280, 1), ("fc2   epi2", M, 1280, 5120, 2),
          ("conv1 epi1", 2 * M, 1280, 384, 1), ("conv2 epi3", M, 1280, 3840, 3), ("plain epi4", M, 1280, 1280, 4)]
g = torch.Generator(device="cuda").manual_seed(0)
for name, m, n, k, epi in shapes:
    A = (torch.randn(m, k, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.zeros(m, n, device="cuda", dtype=torch.half if epi in (0, 1) else torch.float32)
    pos = torch.randn(1500, n, device="cuda", generator=g)
    def run():
        L.check(lib.q2w_op_gemm(A.data_ptr(), k, W.data_ptr(), k, m, n, k, bias.data_ptr(), out.data_ptr(), n, epi,
                                out.data_ptr() if epi == 2 else None, pos.data_ptr(), 1500, n // 2, 0.125, None))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}  M={m} N={n} K={k}: {ms:8.3f} ms  {2.0 * m * n * k / ms / 1e9:8.1f} TFLOP/s")
    del A, W, out
# attention
H, T = 20, 1500
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.5).half()
o = torch.empty(B * T, H * 64, device="cuda", dtype=torch.half)
for nm, fn in (("attention tcgen05", lib.q2w_op_attention),):
    for _ in range(2):
        L.check(fn(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.check(fn(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{nm}: {ms:8.3f} ms  {4.0 * B * T * T * H * 64 / ms / 1e9:8.1f} TFLOP/s")
