#!/bin/bash
# Where does config 1's step (2048x256, single level, fused path through the host classes) go?  (run under gpurun)
H=oracle/_ref/host_harness
run() { echo -n "$1: "; shift; env "$@" 2>&1 | grep ORACLE_TIMING | sed 's/ORACLE_TIMING //'; }
C1="$H /dev/null 2048 256 1 0.1 200 time_only=1 warmup=5"
# (round 2v also timed a 4-deep load ring, VRT_FUSED_NST=4, since removed: no gain — profiles/ab_c1_r2v.txt)
for rep in 1 2; do
run "c1 default" OMP_NUM_THREADS=4 $C1
done
run "c1 Lx=16" OMP_NUM_THREADS=4 VRT_FUSED_LX=16 $C1
run "c1 Lx=64" OMP_NUM_THREADS=4 VRT_FUSED_LX=64 $C1
run "c1 W=96" OMP_NUM_THREADS=4 VRT_FUSED_W=96 $C1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv --log-file gpurun_out/launches_c1.csv $H /dev/null 2048 256 1 0.1 12 pre_steps=0 time_only=1 warmup=2 > gpurun_out/launches_c1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c1.csv "host_harness 2048 256 1 (config 1), plasma phase, 330 launches under ncu"
