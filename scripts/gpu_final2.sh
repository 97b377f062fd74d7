#!/bin/bash
# after the decoder default switch: tests, smoke, bench, ncu of the bench's two kernels
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
CMD="python bench.py --points 40000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD --no-ply > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'encodeTilesKernel|decodePerGaussianKernel' -s 6 -c 2 -o gpurun_out/prof_final -f $CMD --no-ply > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; grep Profiling gpurun_out/ncu_full.log
python scripts/config_sweep.py 2>&1 | tee gpurun_out/configs.jsonl | cut -c1-200
