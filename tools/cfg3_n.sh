#!/bin/bash
# BASELINE configs[3] (Q4_0, 256 windows in total) at N GPUs
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 4 --warmup 3 --wtype q4_0 --total-windows 256 2>gpurun_out/cfg3_n$N.err > gpurun_out/cfg3_q4_0_256_n$N.json
python - "$N" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/cfg3_q4_0_256_n{sys.argv[1]}.json"))
print("cfg3 N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "gemm", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"])
PY
