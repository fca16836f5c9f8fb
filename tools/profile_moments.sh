#!/bin/bash
# ncu --set full of the moments kernel (both species, first stage of the timed step) on C3; run under gpurun
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-self-check --skip-fields-phase --workload ${WL:-c3}"
# warm-up step = 6 stages x 2 species = 12 launches of k_slab_moments
timeout ${TMO:-240} ncu --set full --clock-control none --import-source on -k regex:k_slab_moments -s 12 -c 2 -o gpurun_out/${OUT:-prof_moments} -f $BENCH > gpurun_out/prof_moments_bench.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_moments_bench.log | cut -c1-300
ls -la gpurun_out
