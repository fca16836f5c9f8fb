#!/bin/bash
# Where does the step time of the small AMR workloads go?  (run under gpurun)
H=oracle/_ref/host_harness
run() { echo -n "$1: "; shift; env "$@" 2>&1 | grep ORACLE_TIMING | sed 's/ORACLE_TIMING //'; }
C4="$H /dev/null 512 64 3 0.1 100 time_only=1 warmup=5"
run "c4 regrid22 pdl=1" OMP_NUM_THREADS=4 $C4 regrid_every=22
run "c4 regrid22 pdl=0" OMP_NUM_THREADS=4 VRT_PDL=0 $C4 regrid_every=22
run "c4 noregrid pdl=1" OMP_NUM_THREADS=4 $C4
run "c4 noregrid pdl=0" OMP_NUM_THREADS=4 VRT_PDL=0 $C4
run "c4 noregrid pdl=1 nofork" OMP_NUM_THREADS=4 VRT_HOST_OPT_FORK=0 $C4
C2="$H /dev/null 1024 128 2 0.1 100 refine_mode=1 tail_p0=2 time_only=1 warmup=5"
run "c2 regrid22 pdl=1" OMP_NUM_THREADS=4 $C2 regrid_every=22
run "c2 noregrid pdl=1" OMP_NUM_THREADS=4 $C2
run "c2 noregrid pdl=0" OMP_NUM_THREADS=4 VRT_PDL=0 $C2
# launch list of the plasma phase (no fields phase: pre_steps=0), 3 steps after 3 warm-up steps
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches_c4_plasma.csv $H /dev/null 512 64 3 0.1 8 pre_steps=0 time_only=1 warmup=2 > gpurun_out/launches_c4_plasma.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c4_plasma.csv "host_harness 512 64 3 (config 4), plasma phase, 1200 launches under ncu" | head -40
