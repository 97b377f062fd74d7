#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== default (flat, 3 ctas/sm)"; python scripts/kernel_sweep.py 1e7,1e8 3,0,1,2 2>&1 | cut -c1-110
echo "== 4 ctas/sm (48 regs)"; SPZB200_LIB=spz_b200/_lib/variants/libspz_c4.so python scripts/kernel_sweep.py 1e7,1e8 3,0 2>&1 | cut -c1-110
echo "== 4 ctas/sm persistent"; SPZB200_GRID=persistent SPZB200_CTAS_PER_SM=4 SPZB200_LIB=spz_b200/_lib/variants/libspz_c4.so python scripts/kernel_sweep.py 1e8 3 2>&1 | cut -c1-110
