#!/bin/bash
# split-K A/B at the single-window size + the whole GPU tier
set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for m in 0 1 2; do
  echo "== Q2W_GEMM_SPLITK=$m"
  Q2W_GEMM_SPLITK=$m timeout 300 python tools/gemm_b1.py 2>&1 | tail -8
  Q2W_GEMM_SPLITK=$m timeout 300 python tools/latency_b1.py 2>&1 | head -3
done
