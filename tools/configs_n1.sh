#!/bin/bash
timeout 900 python bench.py --steps 3 --warmup 3 --wtype q4_0 --total-windows 256 --no-cpu-baseline 2>gpurun_out/cfg3_n1.err > gpurun_out/cfg3_q4_0_256_n1.json
tail -3 gpurun_out/cfg3_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/cfg3_q4_0_256_n1.json"))
print("cfg3 N1 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "gemm", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"], d["scaling"], d["config"]["total_windows"])
PY
