#!/bin/bash
# parity first, then kernel sweep under a few launch-shape knobs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== default"; python scripts/kernel_sweep.py 1e7,4e7,1e8 3 2>&1 | cut -c1-110
echo "== flat grid"; SPZB200_GRID=flat python scripts/kernel_sweep.py 1e7,4e7,1e8 3 2>&1 | cut -c1-110
echo "== 2 ctas/sm"; SPZB200_CTAS_PER_SM=2 python scripts/kernel_sweep.py 1e7,1e8 3 2>&1 | cut -c1-110
echo "== 1 cta/sm"; SPZB200_CTAS_PER_SM=1 python scripts/kernel_sweep.py 1e7,1e8 3 2>&1 | cut -c1-110
echo "== other degrees"; python scripts/kernel_sweep.py 1e7,1e8 0,1,2 2>&1 | cut -c1-110
