#!/bin/bash
# round-end sequence on one GPU: smoke(), the whole -m gpu tier, the default bench line, the reference arm
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-200
