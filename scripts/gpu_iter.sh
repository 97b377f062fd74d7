#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 300 python scripts/sanitize_case.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bulk (default)"; timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,2,1 2>&1 | cut -c1-100
echo "== direct"; SPZB200_DECODE=direct timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,2,1 2>&1 | cut -c1-100
