#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "projector or attention or tiny_full or batch_equals or async" 2>&1 | tail -3
timeout 120 python tools/attn_bench.py 64 | tail -1
timeout 120 python tools/attn_bench.py 64 | tail -1
