#!/bin/bash
# whole GPU tier + default bench line (with the parity block and the Q8_0 second pass)
set -x
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "clk", d["clocks"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
print("parity", json.dumps(d["parity"]))
for k, v in d.get("configs", {}).items():
    print(k, "value", round(v["value"]), "e2e", round(v["e2e"]["value"]), "gemm frac", round(v["roofline"]["frac"], 3), "p50_b1", v["p50_ms_per_window_b1"], "parity", json.dumps(v["parity"]))
PY
