#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -2
timeout 300 python tools/gemm_bench.py 64 2>&1 | head -7
timeout 300 python tools/gemm_b1.py 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-wtype > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo rc $?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_t.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50", round(d["p50_ms_per_window_b1"],3), "clk", d["clocks"]["sm_mhz"], "parity", d["parity_ok"])
PY
