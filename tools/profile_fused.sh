#!/bin/bash
# ncu --set full of the fused stage kernel (stage 3, species 0) in the fields-on state; run under gpurun
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-self-check --workload ${WL:-c3}"
# fields phase launches no k_fused_stage; warm-up = 3 steps x 12 launches; take launches of the timed step
ncu --set full --clock-control none --import-source on -k regex:k_fused_stage -s 42 -c 2 -o gpurun_out/${OUT:-prof_fused} -f $BENCH > gpurun_out/prof_bench.log 2>&1
tail -2 gpurun_out/prof_bench.log | cut -c1-300
