#!/bin/bash
export SPZB200_NO_REBUILD=1
timeout 300 python scripts/sanitize_case.py 2>&1 | tail -1
echo "== sweep"; timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,2,1 2>&1 | cut -c1-100
