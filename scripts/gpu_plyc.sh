#!/bin/bash
# canonical-layout PLY kernels + bulk-copy cache hints: device-timed sweeps of the shipped library
# and of the variants under spz_b200/_lib/variants
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
echo "== default";    timeout 300 python scripts/ply_sweep.py 2>&1 | grep 40000000 | cut -c1-140; timeout 300 python scripts/kernel_sweep.py 1e8 3 2>&1 | cut -c1-80
for v in spz_b200/_lib/variants/*.so; do echo "== $v"; SPZB200_LIB=$v timeout 300 python scripts/sanitize_case.py 2>&1 | tail -1; SPZB200_LIB=$v timeout 300 python scripts/ply_sweep.py 2>&1 | grep 40000000 | cut -c1-140;  SPZB200_LIB=$v timeout 300 python scripts/kernel_sweep.py 1e8 3 2>&1 | cut -c1-80; done
