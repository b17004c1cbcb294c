#!/bin/bash
# A/B: TMA L2 prefetch of the streamed A operand, 8 / 16 k-blocks ahead (timeline builds print the MMA thread's wait shares)
cp qwen2_audio_whisper_ggml_b200/libq2w_b200.so /tmp/lib_orig.so
for v in pf8 pf16; do
  cp build_exp/libq2w_$v.so qwen2_audio_whisper_ggml_b200/libq2w_b200.so
  echo "== $v"
  timeout 200 python tools/gemm_bench.py 64 2>&1 | grep -E "TFLOP|^gemm M-tiles 375 N-tiles (5|20|15) K (5120|1280)" | awk '!/^gemm/ || ++n[$4 $6 $8] == 3' | cut -c1-260
done
cp /tmp/lib_orig.so qwen2_audio_whisper_ggml_b200/libq2w_b200.so
