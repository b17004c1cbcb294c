#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "layernorm" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-wtype --no-latency > gpurun_out/bench_ln.json 2> gpurun_out/bench_ln.err; echo rc $?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ln.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "LN GB/s", round(d["kernels"]["layernorm"]["gbs"]), "clk", d["clocks"]["sm_mhz"], "parity", d["parity_ok"])
PY
