#!/bin/bash
for wt in q8_0 q4_0; do
  timeout 900 python bench.py --steps 4 --warmup 3 --wtype $wt --no-cpu-baseline 2>/dev/null > gpurun_out/bench_r1_$wt.json
  python - "$wt" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_r1_{sys.argv[1]}.json"))
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "gemm", round(d["roofline"]["achieved"]), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "setup_s", round(d["setup_s"]))
PY
done
