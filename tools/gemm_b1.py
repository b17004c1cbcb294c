"""Per-shape GEMM / attention / LayerNorm time at the single-window size (M = 1500), the way the B = 1 forward runs them:
each op launched back to back on one stream inside a CUDA graph (so launch overhead is off the clock and PDL chaining is on),
reported as microseconds per launch.  python tools/gemm_b1.py [M]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L

lib = L.load_library()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
g = torch.Generator(device="cuda").manual_seed(0)
st = torch.cuda.Stream()
REPS = 64


def timed(name, fn, flops=0.0):
    with torch.cuda.stream(st):
        for _ in range(3):
            fn(st.cuda_stream)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(REPS):
                fn(st.cuda_stream)
        graph.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(st)
            graph.replay()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) / REPS)
    print(f"{name:28s} {1e3 * best:7.2f} us" + (f"  {flops / best / 1e9:7.1f} TFLOP/s" if flops else ""), flush=True)
    return best


shapes = [("qkv  N=3840 K=1280 epi0", 3840, 1280, 0), ("out  N=1280 K=1280 epi2", 1280, 1280, 2), ("fc1  N=5120 K=1280 epi1", 5120, 1280, 1),
          ("fc2  N=1280 K=5120 epi2", 1280, 5120, 2)]
total = 0.0
for name, n, k, epi in shapes:
    A = (torch.randn(M, k, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.zeros(M, n, device="cuda", dtype=torch.half if epi in (0, 1) else torch.float32)
    total += timed(name, lambda s: L.check(lib.q2w_op_gemm(A.data_ptr(), k, W.data_ptr(), k, M, n, k, bias.data_ptr(), out.data_ptr(), n, epi,
                                                           out.data_ptr() if epi == 2 else None, None, 0, n // 2, 0.125, s)), 2.0 * M * n * k)
B = max(1, M // 1500)
H, T = 20, 1500
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.5).half()
o = torch.empty(B * T, H * 64, device="cuda", dtype=torch.half)
total += timed("attention", lambda s: L.check(lib.q2w_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, s)), 4.0 * B * T * T * H * 64)
x = torch.randn(M, 1280, device="cuda", generator=g)
gam = torch.randn(1280, device="cuda", generator=g)
y = torch.empty(M, 1280, device="cuda", dtype=torch.half)
total += 2 * timed("layernorm", lambda s: L.check(lib.q2w_op_layernorm(x.data_ptr(), gam.data_ptr(), gam.data_ptr(), y.data_ptr(), M, 1280, 1e-5, s)))
print(f"one layer (4 GEMMs + attention + 2 LN), back to back: {1e3 * total:.1f} us; x 32 layers = {32 * total:.3f} ms")
