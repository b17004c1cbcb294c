#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -2
timeout 300 python tools/gemm_b1.py 2>&1 | tail -4
timeout 300 python tools/latency_b1.py 2>&1 | head -5
