#!/bin/bash
# quick tuning sweep of the fused stage kernel on C3 (run under gpurun): prints value and fused-stage throughput per setting
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('value=%.3e e2e=%.3e stage=%.3e frac=%.4f ms/step=%.1f per_stage=%s breakdown=%s clocks=%s'%(d['value'],d['e2e']['value'],r['stage_cell_updates_per_s'],r['frac'],d['ms_per_step'],r['per_stage_GBps'],d.get('breakdown_ms_per_step'),d['clocks']))
except Exception as e: print('ERR',e)
"; }
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg bash -c "$(declare -f run); run $EXTRA"
done
