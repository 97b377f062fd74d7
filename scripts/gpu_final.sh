#!/bin/bash
# round-end evidence on one B200: tests, smoke, both bench arms, ncu launch list and full captures
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
{ nproc; free -g | head -2; nvidia-smi -L; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,clocks_throttle_reasons.active,power.draw,temperature.gpu --format=csv > gpurun_out/clocks_after_bench.csv
CMD="python bench.py --points 40000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD --no-ply > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'Tiles(Bulk)?Kernel' -s 6 -c 2 -o gpurun_out/prof_final -f $CMD --no-ply > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
CMD2="python scripts/ply_prof_target.py 4e7 3"
$CMD2 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'PerGaussian|PlyCanon' -s 2 -c 2 -o gpurun_out/prof_plyc_final -f $CMD2 > gpurun_out/ncu_plyc.log 2>&1
echo "plyc rc=$?"
CMD3="python scripts/prof_sh0_target.py"
$CMD3 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'TilesKernel' -s 3 -c 3 -o gpurun_out/prof_sh0_final -f $CMD3 > gpurun_out/ncu_sh0.log 2>&1
echo "sh0 rc=$?"
