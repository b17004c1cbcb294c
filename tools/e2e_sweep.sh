#!/bin/bash
# e2e (host-buffer) throughput for different copy/compute overlap splits
for sp in 1 2 4; do
  Q2W_E2E_SPLIT=$sp timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>/dev/null > /tmp/b_$sp.json
  python - "$sp" <<'PY'
import json, sys
d = json.load(open(f"/tmp/b_{sys.argv[1]}.json"))
print("split", sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"], 1), "dev_ms", round(d["ms_per_step"], 1), d["clocks"]["sm_mhz"])
PY
done
