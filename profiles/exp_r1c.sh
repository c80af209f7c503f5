#!/bin/bash
# one-call experiment pack: wait-time trace of the producer/consumer backward, no-load variants, clocks under load
mkdir -p gpurun_out
L=tricolo_b200/lib
{
echo "== trace"; TRICOLO_B200_LIB=$L/libtcl_trace.so timeout 120 python profiles/pc_trace.py 8192
echo "== default"; timeout 120 python profiles/time_step.py 8192 50
echo "== exp1 (producer ring never loaded)"; TRICOLO_B200_LIB=$L/libtcl_exp1.so timeout 120 python profiles/time_step.py 8192 50
echo "== exp2 (no ring loads at all)"; TRICOLO_B200_LIB=$L/libtcl_exp2.so timeout 120 python profiles/time_step.py 8192 50
echo "== clocks under a 4 s loop of steps"
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 200 > gpurun_out/exp_clocks.csv &
SMI=$!
timeout 120 python profiles/time_step.py 8192 2000
kill $SMI
sort gpurun_out/exp_clocks.csv | uniq -c | sort -rn | head -12
} > gpurun_out/exp_r1c.log 2>&1
tail -60 gpurun_out/exp_r1c.log
