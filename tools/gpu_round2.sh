#!/bin/bash
# Round-2 GPU session (run under gpurun, 1 GPU): parity tests, smoke, the four bench workloads (both arms for the AMR ones),
# ncu --set full of the fused stage kernel at RK stages 0, 3 and 5 and of the moments kernel, the ncu launch list of a bench step.
#   TAG=r2c KV=v17 bash tools/gpu_round2.sh
# Outputs under gpurun_out/; summaries under gpurun_out/profiles_$TAG/ ready to be copied to profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r2}
KV=${KV:-$(python -c "import re;print(re.search(r'FUSED_KERNEL_VERSION = \"(\w+)\"', open('bench.py').read()).group(1))")}
P=gpurun_out/profiles_$TAG; mkdir -p $P
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed" gpurun_out/pytest_gpu_$TAG.log | tail -2
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
fi
# ncu --set full: one launch each of stages 0, 3, 5 (species 0 and 1 alternate; the 4th step is profiled: 3 warm-up steps x 2
# species x 3 matching stages = 18 launches skipped)
if [ -z "$SKIP_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:k_fused_stageILi[035]E' -s 18 -c 6 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase > gpurun_out/prof_bench_$TAG.log 2>&1
echo "ncu fused rc=$?"; tail -1 gpurun_out/prof_bench_$TAG.log | cut -c1-160
python tools/ncu_summary.py gpurun_out/prof_fused_$TAG.ncu-rep > $P/ncu_fused_${KV}_${TAG}_summary.txt
python tools/ncu_source_summary.py gpurun_out/prof_fused_$TAG.ncu-rep 30 > $P/ncu_fused_${KV}_${TAG}_source.txt 2>/dev/null
python tools/update_traffic.py $KV profiles/ncu_fused_${KV}_${TAG}_summary.txt gpurun_out/prof_fused_$TAG.ncu-rep > /dev/null && cp profiles/fused_traffic.json $P/
OUT=prof_moments_$TAG bash tools/profile_moments.sh > /dev/null
python tools/ncu_summary.py gpurun_out/prof_moments_$TAG.ncu-rep > $P/ncu_moments_${TAG}_summary.txt
python tools/ncu_source_summary.py gpurun_out/prof_moments_$TAG.ncu-rep 30 > $P/ncu_moments_${TAG}_source.txt 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase > gpurun_out/launches_bench_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-self-check --skip-fields-phase under ncu --metrics gpu__time_duration.sum ($KV)" > $P/launches_${TAG}_${KV}_summary.txt 2>/dev/null
head -8 $P/launches_${TAG}_${KV}_summary.txt
fi
# launch list of the 3-level AMR workload through the host classes (what the step is made of)
# (pre_steps=0: no fields-only phase, so that the window lands in the Vlasov steps)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches_c4_$TAG.csv oracle/_ref/host_harness /dev/null 512 64 3 0.1 8 pre_steps=0 time_only=1 warmup=2 > gpurun_out/launches_c4_bench_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c4_$TAG.csv "host_harness 512 64 3 (config 4), plasma phase, 1200 launches under ncu --metrics gpu__time_duration.sum" > $P/launches_c4_${TAG}_summary.txt 2>/dev/null
head -40 $P/launches_c4_${TAG}_summary.txt
for wl in c1 c2 c4; do
  timeout 600 python bench.py --workload $wl --steps 100 --warmup 5 > $P/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; cut -c1-400 $P/bench_${wl}_$TAG.json
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $P/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-300 $P/bench_ref_$TAG.json
timeout 900 python bench.py > $P/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-3000 $P/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
