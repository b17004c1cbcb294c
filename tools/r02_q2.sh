#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "tiny_full or tiny_quant or full_size_vs_golden" 2>&1 | tail -2
Q2W_ONLY=0 bash tools/r02_quant_b1.sh 2>&1 | grep "fused=0"
