#!/bin/bash
# Short-grid Poisson solve: cluster of CTAs (mode 2) against one wide CTA with a thread group per tile (mode 3).  (run under gpurun)
H=oracle/_ref/host_harness
run() { echo -n "$1: "; shift; env "$@" 2>&1 | grep ORACLE_TIMING | sed 's/ORACLE_TIMING //'; }
C1="$H /dev/null 2048 256 1 0.1 200 time_only=1 warmup=5"
C2="$H /dev/null 1024 128 2 0.1 100 refine_mode=1 tail_p0=2 time_only=1 warmup=5 regrid_every=22"
C4="$H /dev/null 512 64 3 0.1 100 time_only=1 warmup=5 regrid_every=22"
for m in 2 3 2 3; do run "c1 poisson mode $m" OMP_NUM_THREADS=4 VRT_POISSON_SMALL=$m $C1; done
for m in 2 3; do run "c2 poisson mode $m" OMP_NUM_THREADS=4 VRT_POISSON_SMALL=$m $C2; done
for m in 2 3 2 3; do run "c4 poisson mode $m" OMP_NUM_THREADS=4 VRT_POISSON_SMALL=$m $C4; done
VRT_POISSON_SMALL=3 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_poisson -s 40 -c 12 --csv --log-file gpurun_out/launches_poisson_wide.csv $H /dev/null 2048 256 1 0.1 6 pre_steps=0 time_only=1 warmup=2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_poisson_wide.csv "k_poisson_wide, N = 2048, under ncu"
