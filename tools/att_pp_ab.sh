#!/bin/bash
# ping-pong attention kernel: correctness (kernel tests) + per-layer time at B = 64, against the two-CTAs-per-SM kernel
for pp in 1 0; do
  echo "== Q2W_ATT_PINGPONG=$pp"
  Q2W_ATT_PINGPONG=$pp timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3
  Q2W_ATT_PINGPONG=$pp timeout 120 python tools/attn_bench.py 64
  Q2W_ATT_PINGPONG=$pp timeout 120 python tools/attn_bench.py 64
  Q2W_ATT_PINGPONG=$pp timeout 120 python tools/attn_bench.py 1
done
