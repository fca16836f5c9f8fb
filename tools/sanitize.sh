#!/bin/bash
# compute-sanitizer over the small GPU tests that reach every kernel family (run under gpurun); summary -> gpurun_out/sanitizer_$TAG.txt
TAG=${TAG:-r2}
OUT=gpurun_out/sanitizer_$TAG.txt; mkdir -p gpurun_out; : > $OUT
SEL='test_stage_parity_with_injected_phi or test_per_step_parity_from_reference_state or test_poisson_single_cta or test_fused_equals_split_ragged or test_moment_kernel_variants_short or test_amr_phases_against_oracle or test_amr_per_step or test_device_initial_condition or test_checkpoint'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool, pytest -k '$SEL'" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_amr.py tests/test_gpu_initial_condition.py tests/test_gpu_checkpoint.py -q -x -k "$SEL" > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_$TAG.log | tail -4 >> $OUT
  grep -E "Invalid|hazard|Race reported" gpurun_out/sanitizer_${tool}_$TAG.log | sort | uniq -c | head -10 >> $OUT
done
cat $OUT
