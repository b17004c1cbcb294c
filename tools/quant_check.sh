#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -k "q8_0 or q4_0 or quant or dequant or custom_loader or hf_qwen2 or attention" 2>&1 | tail -3
bash tools/quant_b1.sh 2>&1 | grep "fused=0"
timeout 120 python tools/attn_bench.py 64 | tail -1
timeout 600 python bench.py --wtype q8_0 --steps 5 --warmup 3 --no-cpu-baseline --no-second-wtype > gpurun_out/bench_q8.json 2> gpurun_out/bench_q8.err; echo rc $?; tail -2 gpurun_out/bench_q8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q8.json"))
print("q8_0 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm frac", round(d["roofline"]["frac"], 3), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "parity_ok", d["parity_ok"])
PY
