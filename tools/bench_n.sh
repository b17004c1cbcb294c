#!/bin/bash
# default bench line at N GPUs (what the driver's scaling run launches)
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json
python - "$N" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_n{sys.argv[1]}.json"))
print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"]), "gather_ms", d["nccl_gather_ms"], "clk", d["clocks"]["sm_mhz"])
PY
