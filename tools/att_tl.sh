#!/bin/bash
cd qwen2_audio_whisper_ggml_b200/csrc && make clean >/dev/null && make -j8 EXTRA_NVFLAGS=-DQ2W_ATT_TIMELINE > /dev/null 2>&1; cd ../..
timeout 300 python tools/att_tl.py ${1:-64} 2>&1 | grep "^g\|^cta 5 \|^cta 153 " | tail -44
