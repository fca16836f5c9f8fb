#!/bin/bash
# Round-1 profiling pass (run under gpurun, 1 GPU): fp64 peak, ncu launch list of one bench run, ncu --set full of the
# fused stage kernel.  Outputs under gpurun_out/; summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
./tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1
cat gpurun_out/fp64_peak.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-fields-phase"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv $BENCH > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log
# stage kernels of the timed region: skip init + warm-up launches of k_fused_stage (3 warm-up steps x 12 + ...), take 3
ncu --set full --clock-control none --import-source on -k regex:k_fused_stage -s 40 -c 3 -o gpurun_out/prof_fused_r1 -f $BENCH > gpurun_out/prof_bench.log 2>&1
tail -2 gpurun_out/prof_bench.log
ls -la gpurun_out
