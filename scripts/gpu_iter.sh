#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 300 python scripts/sanitize_case.py 2>&1 | tail -2
echo "== hoist (default)"; timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,1 2>&1 | cut -c1-100
echo "== no hoist"; SPZB200_LIB=spz_b200/_lib/variants/libspz_nohoist.so timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,1 2>&1 | cut -c1-100
