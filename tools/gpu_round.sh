#!/bin/bash
# One GPU session (run under gpurun, 1 GPU): parity tests, smoke, default bench (both arms), ncu launch list, ncu --set full of
# the fused stage kernel.  Outputs under gpurun_out/; summaries are copied to profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-fields-phase"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_bench_$TAG.log 2>&1
tail -1 gpurun_out/launches_bench_$TAG.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_stage -s 42 -c 2 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-fields-phase > gpurun_out/prof_bench_$TAG.log 2>&1
tail -1 gpurun_out/prof_bench_$TAG.log | cut -c1-200
ls -la gpurun_out
