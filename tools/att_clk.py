"""attention at the bench shape in a 3 s loop while sampling SM clock and power (is the kernel power-capped?)"""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen2_audio_whisper_ggml_b200 import lib as L
lib = L.load_library()
B, H, T = 64, 20, 1500
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.5).half()
o = torch.empty(B * T, H * 64, device="cuda", dtype=torch.half)
samples = []
stop = False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        samples.append(r.stdout.strip())
        time.sleep(0.2)
th = threading.Thread(target=sampler); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 2500
e0.record()
for _ in range(n):
    L.check(lib.q2w_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, None))
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
print(f"attention B={B}: {e0.elapsed_time(e1) / n:.3f} ms over {n} launches")
print("clock/power samples:", samples[2:-1][:12])
