#!/bin/bash
# One GPU session (run under gpurun, 1 GPU): parity tests, smoke, default bench (both arms), ncu launch list, ncu --set full of
# the fused stage kernel (with the DRAM traffic written to profiles/fused_traffic.json for bench.py) and of the moments kernel.
# Outputs under gpurun_out/; the summaries are also written to gpurun_out/profiles_$TAG/ ready to be copied to profiles/.
#   TAG=r2a KV=v16 bash tools/gpu_round.sh        (KV = FUSED_KERNEL_VERSION of bench.py)
mkdir -p gpurun_out
TAG=${TAG:-r1}
KV=${KV:-$(python -c "import re;print(re.search(r'FUSED_KERNEL_VERSION = \"(\w+)\"', open('bench.py').read()).group(1))")}
P=gpurun_out/profiles_$TAG; mkdir -p $P
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_$TAG.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-fields-phase"
# ncu passes first: the full capture's DRAM traffic goes into the bench line of this session
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_stage -s 42 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-fields-phase > gpurun_out/prof_bench_$TAG.log 2>&1
tail -1 gpurun_out/prof_bench_$TAG.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/prof_fused_$TAG.ncu-rep > $P/ncu_fused_${KV}_${TAG}_summary.txt
python tools/ncu_source_summary.py gpurun_out/prof_fused_$TAG.ncu-rep 30 > $P/ncu_fused_${KV}_${TAG}_source.txt
python tools/update_traffic.py gpurun_out/prof_fused_$TAG.ncu-rep $KV profiles/ncu_fused_${KV}_${TAG}_summary.txt && cp profiles/fused_traffic.json $P/
OUT=prof_moments_$TAG bash tools/profile_moments.sh > /dev/null
python tools/ncu_summary.py gpurun_out/prof_moments_$TAG.ncu-rep > $P/ncu_moments_${TAG}_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_bench_$TAG.log 2>&1
tail -1 gpurun_out/launches_bench_$TAG.log | cut -c1-200
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > $P/launches_${TAG}_${KV}_summary.txt 2>/dev/null
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $P/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 $P/bench_ref_$TAG.json
timeout 900 python bench.py > $P/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 $P/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
ls -la gpurun_out $P
