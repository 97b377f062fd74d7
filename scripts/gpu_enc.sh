#!/bin/bash
# planar encoder: bulk-copy per-gaussian kernel (default) vs register-path tiles (SPZB200_ENCODE=direct)
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== default";    timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,0,1,2 2>&1 | cut -c1-80
echo "== direct";     SPZB200_ENCODE=direct timeout 300 python scripts/kernel_sweep.py 1e7,1e8 3,0,1,2 2>&1 | cut -c1-80
echo "== persistent"; SPZB200_GRID=persistent timeout 300 python scripts/kernel_sweep.py 1e8 3,0 2>&1 | cut -c1-80
