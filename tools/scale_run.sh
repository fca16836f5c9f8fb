#!/bin/bash
# strong-scaling bench lines on one multi-GPU box (run under gpurun --gpus 8): N in "$@" (default 4 8)
mkdir -p gpurun_out
for N in ${@:-4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.log 2>&1
  grep '^{"metric"' gpurun_out/scale_n$N.log > gpurun_out/scale_n$N.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print("N=$N value=%.4e e2e=%.4e ms/step=%.2f frac=%.4f breakdown=%s checks=%s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["breakdown_ms_per_step"], d["checks"]))
PY
done
