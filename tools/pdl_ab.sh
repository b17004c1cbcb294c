#!/bin/bash
# A/B of programmatic dependent launch: correctness tier first, then B=1 latency and the default bench with PDL on / off
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for pdl in 1 0; do
  echo "== Q2W_PDL=$pdl"
  Q2W_PDL=$pdl timeout 300 python tools/latency_b1.py 1 2>&1 | grep -E "p50|total|gemm|attention" | head -6
  Q2W_PDL=$pdl timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_pdl$pdl.json
  python - "$pdl" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_pdl{sys.argv[1]}.json"))
print("bench value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "p50_b1", d.get("latency_b1_ms_p50"), "clk", d["clocks"]["sm_mhz"])
PY
done
