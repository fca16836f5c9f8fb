#!/bin/bash
# What does updating the parked step-graph executable after a regrid (instead of instantiating a new one) buy?  (run under gpurun)
H=oracle/_ref/host_harness
run() { echo -n "$1: "; shift; env "$@" 2>&1 | grep ORACLE_TIMING | sed 's/ORACLE_TIMING //'; }
C4="$H /dev/null 512 64 3 0.1 100 time_only=1 warmup=5"
C2="$H /dev/null 1024 128 2 0.1 100 refine_mode=1 tail_p0=2 time_only=1 warmup=5"
for u in 1 0 1 0; do
run "c4 regrid22 update=$u" OMP_NUM_THREADS=4 VRT_GRAPH_UPDATE=$u $C4 regrid_every=22
done
run "c4 noregrid" OMP_NUM_THREADS=4 $C4
for u in 1 0; do
run "c2 regrid22 update=$u" OMP_NUM_THREADS=4 VRT_GRAPH_UPDATE=$u $C2 regrid_every=22
done
