#!/bin/bash
mkdir -p gpurun_out
python scripts/kernel_sweep.py 1.25e6,2.5e6,5e6,1e7,2e7,4e7,1e8 3 > gpurun_out/sweep_sh3.jsonl 2>&1
python scripts/kernel_sweep.py 1e7,1e8 0,1,2 > gpurun_out/sweep_sh012.jsonl 2>&1
cat gpurun_out/sweep_sh3.jsonl gpurun_out/sweep_sh012.jsonl
CMD="python bench.py --points 20000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'encode|decode' -c 40 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'TilesKernel' -s 6 -c 2 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
