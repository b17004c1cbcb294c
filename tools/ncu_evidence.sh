#!/bin/bash
# round-2 ncu evidence for the CURRENT kernels: four per-layer GEMM shapes at M = 96000, mel, launch lists at B = 64 and B = 1
set -x
O=gpurun_out
for s in "qkv 3840 1280 0" "outproj 1280 1280 2" "fc1 5120 1280 1" "fc2 1280 5120 2"; do
  set -- $s
  python tools/gemm_one.py 96000 $2 $3 $4 > $O/plain_gemm_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o $O/r02_gemm_$1 python tools/gemm_one.py 96000 $2 $3 $4 > $O/ncu_gemm_$1.log 2>&1
  tail -1 $O/ncu_gemm_$1.log
done
python tools/mel_one.py 64 > $O/plain_mel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mel_logpower -s 2 -c 1 -o $O/r02_mel python tools/mel_one.py 64 > $O/ncu_mel.log 2>&1
tail -1 $O/ncu_mel.log; cat $O/plain_mel.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-second-wtype --no-latency > $O/bench_for_ncu.json 2> $O/bench_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 236 -c 470 --csv --log-file $O/r02_launches_bench_default_b64.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-second-wtype --no-latency > $O/ncu_bench.log 2>&1
tail -2 $O/ncu_bench.log
python tools/one_window.py > $O/plain_one.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file $O/r02_launches_b1.csv python tools/one_window.py > $O/ncu_one.log 2>&1
tail -2 $O/ncu_one.log
python tools/attn_bench.py 64 > $O/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o $O/r02_attn_final python tools/attn_bench.py 64 > $O/ncu_attn.log 2>&1
tail -1 $O/ncu_attn.log
