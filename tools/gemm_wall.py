"""Drive the -DQ2W_GEMM_WALL diagnostic build: a short chain out-proj -> fc1-shaped -> fc2 at M = 1500 inside a CUDA graph; the
kernels print globaltimer stamps (ns) for their first and last cluster; this script turns them into a per-launch timeline.
  make -C qwen2_audio_whisper_ggml_b200/csrc clean && make -C ... EXTRA_NVFLAGS=-DQ2W_GEMM_WALL && python tools/gemm_wall.py"""
import sys, os, re, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from qwen2_audio_whisper_ggml_b200 import lib as L
    lib = L.load_library()
    M = 1500
    g = torch.Generator(device="cuda").manual_seed(0)
    st = torch.cuda.Stream()
    shapes = [(1280, 1280, 2), (5120, 1280, 1), (1280, 5120, 2), (3840, 1280, 0)]
    ops = []
    for n, k, epi in shapes:
        A = (torch.randn(M, k, device="cuda", generator=g) * 0.5).half()
        W = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).half()
        bias = torch.randn(n, device="cuda", generator=g)
        out = torch.zeros(M, n, device="cuda", dtype=torch.half if epi in (0, 1) else torch.float32)
        ops.append((A, W, bias, out, n, k, epi))
    def chain(s):
        for A, W, bias, out, n, k, epi in ops:
            L.check(lib.q2w_op_gemm(A.data_ptr(), k, W.data_ptr(), k, M, n, k, bias.data_ptr(), out.data_ptr(), n, epi,
                                    out.data_ptr() if epi == 2 else None, None, 0, n // 2, 0.125, s))
    with torch.cuda.stream(st):
        chain(st.cuda_stream); st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(3):
                chain(st.cuda_stream)
        print("==GRAPH", flush=True)
        gr.replay(); st.synchronize()
        print("==REPLAY2", flush=True)
        gr.replay(); st.synchronize()
    sys.exit(0)
out = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True).stdout
seg = out.split("==REPLAY2")[-1]
ev = {}
order = []
for ln in seg.splitlines():
    m = re.match(r"GW (\w+) c(\d+)/(\d+) N(\d+) K(\d+) (.*)", ln)
    if not m:
        continue
    kind, c, nc, N, K, rest = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)), m.group(6)
    vals = dict(zip(rest.split()[0::2], [int(x) for x in rest.split()[1::2]]))
    ev.setdefault((N, K, c, nc), []).append((kind, vals))
# launches are serialised by the stream: group stamps by time
recs = []
for (N, K, c, nc), lst in ev.items():
    mm = [v for k, v in lst if k == "mma"]
    ep = [v for k, v in lst if k == "epi"]
    en = [v for k, v in lst if k == "end"]
    for i in range(min(len(mm), len(ep), len(en))):
        recs.append((mm[i]["entry"], N, K, c, nc, mm[i], ep[i], en[i]))
recs.sort()
t0 = recs[0][0] if recs else 0
prev_exit = None
for entry, N, K, c, nc, mm, ep, en in recs:
    f = lambda x: (x - t0) / 1e3
    print(f"N{N:5d} K{K:5d} cluster {c:2d}/{nc:2d}: entry {f(entry):8.2f} prolog +{(mm['prolog'] - entry) / 1e3:5.2f} wait_done {f(mm['wait']):8.2f} first_operands +{(mm['first_operands'] - mm['wait']) / 1e3:5.2f} "
          f"mma_issued +{(mm['issued_all'] - mm['first_operands']) / 1e3:5.2f} acc_ready +{(ep['acc_ready'] - mm['issued_all']) / 1e3:5.2f} stores_issued +{(ep['stores_issued'] - ep['acc_ready']) / 1e3:5.2f} "
          f"stores_read +{(ep['stores_read'] - ep['stores_issued']) / 1e3:5.2f} exit {f(en['exit']):8.2f} (+{(en['exit'] - ep['stores_read']) / 1e3:5.2f})")
