#!/bin/bash
# Drop-in comparison on the GPU box: ONE case file (oracle/ref_harness.cpp) built against the unmodified reference classes
# (oracle/_ref/ref_harness, CPU, all host cores) and against the veritas_b200 host classes (oracle/_ref/host_harness, B200),
# run on BASELINE.json configs 1, 2 and 4 (the ones the reference can run).  Timed: CalculateDt + Advance of every step.
export OPENBLAS_NUM_THREADS=1
STEPS=${STEPS:-30}
run() {  # label, args...
  local label=$1; shift
  for exe in ref_harness host_harness; do
    local line=$(OMP_NUM_THREADS=$(nproc) timeout 900 oracle/_ref/$exe /dev/null "$@" time_only=1 2>&1 | grep ORACLE_TIMING | tail -1)
    echo "$label | $exe | $line"
  done
}
run "config 1: 2048x256 single level"            2048 256 1 0.1 $STEPS
run "config 2: 1024x128 coarse, 2 levels, tail"  1024 128 2 0.1 $STEPS refine_mode=1 tail_p0=2
run "config 4: 512x64 coarse, 3 levels, regrid 22" 512 64 3 0.1 $STEPS regrid_every=22
