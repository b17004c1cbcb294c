#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -k "full_size_vs_golden or tiny_full" 2>&1 | tail -3
Q2W_NEXT_W_PREFETCH=0 timeout 300 python tools/latency_b1.py 2>&1 | head -3
Q2W_NEXT_W_PREFETCH=1 timeout 300 python tools/latency_b1.py 2>&1 | head -3
Q2W_NEXT_W_PREFETCH=0 timeout 300 python tools/latency_b1.py 2>&1 | head -1
Q2W_NEXT_W_PREFETCH=1 timeout 300 python tools/latency_b1.py 2>&1 | head -1
