#!/bin/bash
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_exp.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_exp.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "clk", d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
PY
