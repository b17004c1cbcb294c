#!/bin/bash
# A/B of the FMA-pipe exponential share in the attention softmax: correctness (kernel tests) + per-layer time at B = 64
for np in 0 1 2 3 4; do
  echo "== Q2W_ATT_POLY=$np"
  Q2W_ATT_POLY=$np timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -1
  Q2W_ATT_POLY=$np timeout 120 python tools/attn_bench.py 64
  Q2W_ATT_POLY=$np timeout 120 python tools/attn_bench.py 64
done
