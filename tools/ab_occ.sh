#!/bin/bash
# does the stage kernel need 4 CTAs per SM in its HBM-bound stages?  (run under gpurun)
for pad in "" "3:16,4:16,5:16" "3:32,4:32,5:32" "0:16,1:16,2:16"; do
  VRT_FUSED_PAD_KB=$pad python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase > gpurun_out/ab_occ.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_occ.json'));print('PAD=$pad','ms/step %.2f'%d['ms_per_step'],d['roofline']['frac'],d['roofline']['per_stage_GBps'])"
done
