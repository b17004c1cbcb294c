#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -m gpu -q -x -k "gemm or full_size_vs_golden or tiny_full or five_clips" 2>&1 | tail -4
timeout 300 python tools/gemm_bench.py 64 2>&1 | head -8
timeout 300 python tools/gemm_b1.py 2>&1 | tail -8
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-second-wtype --no-latency > $O/bench_for_ncu.json 2> $O/bench_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 236 -c 470 --csv --log-file $O/r02_launches_bench_default_b64.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-second-wtype --no-latency > $O/ncu_bench.log 2>&1
tail -2 $O/ncu_bench.log | cut -c 1-300
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()}, "p50_b1", round(d["p50_ms_per_window_b1"], 3), "clk", d["clocks"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
print("traffic", d["roofline"]["traffic"], d["roofline"]["traffic_algorithmic"], "parity_ok", d["parity_ok"], {k: round(v["rel_l2"], 6) for k, v in d["parity"].items()})
for k, v in d.get("configs", {}).items():
    print(k, "value", round(v["value"]), "e2e", round(v["e2e"]["value"]), "gemm frac", round(v["roofline"]["frac"], 3), "p50_b1", v["p50_ms_per_window_b1"], {kk: round(vv["rel_l2"], 6) for kk, vv in v["parity"].items()})
PY
