#!/bin/bash
# BASELINE configs[2], [3]@8 and [4] on one 8-GPU box (strong scaling: the total number of windows is fixed)
N=${1:-8}
run() {  # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 4 --warmup 3 "$@" 2>gpurun_out/${name}_n$N.err > gpurun_out/${name}_n$N.json
  python - "$name" "$N" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_n{sys.argv[2]}.json"))
    print(sys.argv[1], "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "gemm", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"], "setup_s", round(d["setup_s"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run cfg2_q8_0_256 --wtype q8_0 --total-windows 256
run cfg3_q4_0_256 --wtype q4_0 --total-windows 256
run cfg4_long1h_f16 --wtype f16 --total-windows 120
