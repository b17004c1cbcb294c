#!/bin/bash
# whole GPU tier, single-window breakdown, default bench line
set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python tools/gemm_b1.py 2>&1 | tail -8
timeout 300 python tools/latency_b1.py 2>&1 | head -8
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "clk", d["clocks"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
PY
