#!/bin/bash
# A/B of fused-stage kernel variants in one GPU session (run under gpurun): parity first, then short device-resident benches.
#   MASKS="0x00 0x0f 0x3f" bash tools/ab_fused.sh
mkdir -p gpurun_out
for m in ${MASKS:-0x00 0x0f 0x3f}; do
  export VRT_FUSED_LEAN=$m
  if [ "$m" != "0x00" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_on_box.py tests/test_gpu_checkpoint.py -q -k "fused or ragged or interior or benchmark_p_grid or hundred or checkpoint or per_step" > gpurun_out/ab_pytest_$m.log 2>&1
    echo "LEAN=$m pytest rc=$? $(tail -1 gpurun_out/ab_pytest_$m.log)"
  fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase > gpurun_out/ab_bench_$m.json 2> gpurun_out/ab_bench_$m.err
  python - <<P
import json
d=json.load(open("gpurun_out/ab_bench_$m.json"))
print("LEAN=$m", "ms/step %.2f" % d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["per_stage_GBps"], d["breakdown_ms_per_step"])
P
done
