#!/bin/bash
# A/B of the fused stage kernel's x-chunk length and graded tail (run under gpurun): short device-resident benches on config 3
mkdir -p gpurun_out
for cfg in "256 64" "384 64" "384 96" "512 64" "512 128" "192 48"; do
  set -- $cfg
  VRT_FUSED_LX=$1 VRT_FUSED_TAIL=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase > gpurun_out/ab_chunk_$1_$2.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_chunk_$1_$2.json'));print('LX=$1 TAIL=$2','ms/step %.2f'%d['ms_per_step'],d['roofline']['frac'],d['roofline']['per_stage_GBps'],d['breakdown_ms_per_step']['fused_stage'])"
done
