#!/bin/bash
# ncu evidence of round 2 at HEAD (one GPU): launch list of the bench command + full captures of the loss and retrieval kernels
tag=${1:-r2e}
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-small-batch --retrieval-queries 151552 > gpurun_out/${tag}_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ntxent_|l2norm_|fwd_reduce|fwd_finalize" -s 14 -c 7 -o gpurun_out/${tag}_loss -f \
    python profiles/prof_step.py loss > gpurun_out/${tag}_loss.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sim_topk_fused" -s 1 -c 1 -o gpurun_out/${tag}_fused -f \
    python profiles/prof_step.py fused > gpurun_out/${tag}_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sim_gemm_resident|topk_rank" -c 2 -o gpurun_out/${tag}_twokernel -f \
    python profiles/prof_step.py retrieval > gpurun_out/${tag}_twokernel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"topk_rank" -c 1 -o gpurun_out/${tag}_shard -f \
    python profiles/prof_step.py shard > gpurun_out/${tag}_shard.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ntxent_small" -s 2 -c 2 -o gpurun_out/${tag}_small -f \
    python profiles/prof_step.py loss 256 > gpurun_out/${tag}_small.log 2>&1
tail -2 gpurun_out/${tag}_small.log gpurun_out/${tag}_loss.log gpurun_out/${tag}_fused.log gpurun_out/${tag}_twokernel.log gpurun_out/${tag}_shard.log
ls -la gpurun_out/${tag}_*
