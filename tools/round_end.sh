#!/bin/bash
# the round-end sequence: whole GPU tier, smoke, default bench line (F16 + ride-along configs), reference arm
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0; timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $? in $SECONDS s"; tail -3 $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "clk", d["clocks"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
print("parity_ok", d["parity_ok"], {k: round(v["rel_l2"], 6) for k, v in d["parity"].items() if v.get("checked")})
for k, v in d.get("configs", {}).items():
    print(k, "value", round(v["value"]), "e2e", round(v["e2e"]["value"]), "ms", round(v.get("ms_per_step", v.get("ms_per_pass")), 1), "p50_b1", v.get("p50_ms_per_window_b1"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c 1-400
