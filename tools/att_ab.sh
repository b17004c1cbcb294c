#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3
timeout 120 python tools/attn_bench.py 64
timeout 120 python tools/attn_bench.py 64
timeout 120 python tools/attn_bench.py 1
timeout 120 python tools/attn_bench.py 3
