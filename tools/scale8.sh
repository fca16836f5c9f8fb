#!/bin/bash
# 8-GPU session (run under gpurun --gpus 8): bitwise x-slab checks at 2, 4, 8 ranks, then the bench at N = 8 and N = 4 with default
# arguments (their state hashes must agree).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | tail -6
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) bench.py --gpus $n > gpurun_out/bench_n${n}_r2r.raw 2> gpurun_out/bench_n${n}_r2r.err
  echo "bench n=$n rc=$?"; grep '^{' gpurun_out/bench_n${n}_r2r.raw > gpurun_out/bench_n${n}_r2r.json
  python -c "
import json;d=json.load(open('gpurun_out/bench_n${n}_r2r.json'));print('N=$n',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']['frac'],d['breakdown_ms_per_step'],d['checks']['state_hash'],d['clocks'])"
done
