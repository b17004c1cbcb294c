#!/bin/bash
# A/B of the single-window latency between two builds of the library: put them in build_exp/lib_reduce.so (default build) and
# build_exp/lib_load.so (make EXTRA_NVFLAGS=-DQ2W_GEMM_RESID_LOAD) first; build_exp/ is git-ignored but travels with gpurun
cp qwen2_audio_whisper_ggml_b200/libq2w_b200.so /tmp/lib_orig.so
for v in reduce load reduce load; do
  cp build_exp/lib_$v.so qwen2_audio_whisper_ggml_b200/libq2w_b200.so
  echo "== $v"; timeout 200 python tools/latency_b1.py 1 2>&1 | grep -E "p50|gemm|attention|layernorm" | head -4
done
cp /tmp/lib_orig.so qwen2_audio_whisper_ggml_b200/libq2w_b200.so
