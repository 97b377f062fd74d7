#!/bin/bash
# canonical-layout PLY kernels: parity tests, then device-timed sweeps of the shipped library, the
# column-map kernels (SPZB200_PLY=mapped), multi-tile CTAs and the CTAs/SM variants
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ply or alternate" 2>&1 | tail -5
echo "== default";    timeout 300 python scripts/ply_sweep.py 2>&1 | cut -c1-140
echo "== mapped";     SPZB200_PLY=mapped timeout 300 python scripts/ply_sweep.py 2>&1 | cut -c1-140
echo "== persistent"; SPZB200_GRID=persistent timeout 300 python scripts/ply_sweep.py 2>&1 | cut -c1-140
for v in spz_b200/_lib/variants/*.so; do echo "== $v"; SPZB200_LIB=$v timeout 300 python scripts/ply_sweep.py 2>&1 | cut -c1-140; done
