#!/bin/bash
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -k "async or batch" 2>&1 | tail -2
for sync in 0 1; do
  Q2W_BENCH_E2E_SYNC=$sync timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_e2e_sync$sync.json
  python - "$sync" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_e2e_sync{sys.argv[1]}.json"))
print("sync" if sys.argv[1] == "1" else "async", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "dev ms", round(d["ms_per_step"], 2), "clk", d["clocks"]["sm_mhz"])
PY
done
