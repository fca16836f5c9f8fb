#!/bin/bash
# Round-end confirmation of the committed tree on one B200: GPU test suite, smoke(), the default bench line and the host-class workloads.
TAG=${1:-r2v}
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 300 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; cut -c1-400 $O/bench_$TAG.json
for w in c1 c2 c4; do timeout 200 python bench.py --workload $w > $O/bench_${w}_$TAG.json 2> $O/bench_${w}_$TAG.err; python -c "
import json,sys
d=json.load(open('$O/bench_${w}_$TAG.json')); print('$w', d['value'], d['ms_per_step'], d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)"; done
