#!/bin/bash
# round-2 baseline with the round-1 kernels: B = 1 breakdown, ncu launch list at B = 1, ncu full capture of the current attention kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python tools/gemm_b1.py > gpurun_out/r02_base_gemm_b1.log 2>&1; tail -12 gpurun_out/r02_base_gemm_b1.log
timeout 300 python tools/latency_b1.py > gpurun_out/r02_base_latency_b1.log 2>&1; cat gpurun_out/r02_base_latency_b1.log
timeout 300 python tools/one_window.py > gpurun_out/plain_one.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file gpurun_out/r02_base_launches_b1.csv python tools/one_window.py > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/ncu_one.log
timeout 300 python tools/attn_bench.py 64 > gpurun_out/plain_attn.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o gpurun_out/r02_attn_pulled python tools/attn_bench.py 64 > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log; cat gpurun_out/plain_attn.log
