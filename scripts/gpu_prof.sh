#!/bin/bash
# usage: gpu_prof.sh <round-tag>   -> bench (full) + ncu launch list + ncu --set full of both tile kernels
TAG=${1:-r1}
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_${TAG}.csv
CMD="python bench.py --points 40000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'encode|decode' -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"Tiles.*Kernel" -s 6 -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
