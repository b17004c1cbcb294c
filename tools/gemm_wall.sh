#!/bin/bash
set -x
cd qwen2_audio_whisper_ggml_b200/csrc && make clean >/dev/null && make -j8 EXTRA_NVFLAGS=-DQ2W_GEMM_WALL > /dev/null 2>&1; cd ../..
timeout 300 python tools/gemm_wall.py child 2>&1 | grep "^GC" | tail -24
