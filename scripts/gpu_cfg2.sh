#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 600 python scripts/config_sweep.py 2>&1 | tee gpurun_out/configs.jsonl
CMD="python scripts/prof_sh0_target.py"
$CMD && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'TilesKernel' -s 3 -c 3 -o gpurun_out/prof_sh0 -f $CMD > gpurun_out/ncu_sh0.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_sh0.log
