#!/bin/bash
# one-call experiment pack: wait-time trace of the producer/consumer backward with and without its operand loads
mkdir -p gpurun_out
L=tricolo_b200/lib
{
for v in trace exp1 exp2; do
echo "== $v"; TRICOLO_B200_LIB=$L/libtcl_$v.so timeout 120 python profiles/pc_trace.py 8192
TRICOLO_B200_LIB=$L/libtcl_$v.so timeout 120 python profiles/time_step.py 8192 50
done
echo "== default"; timeout 120 python profiles/time_step.py 8192 50
} > gpurun_out/exp_r1c.log 2>&1
cat gpurun_out/exp_r1c.log
