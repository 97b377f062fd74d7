#!/bin/bash
export SPZB200_NO_REBUILD=1
echo "== 4 CTAs/SM"; python scripts/ply_sweep.py 2>&1 | grep '40000000'
for v in spz_b200/_lib/variants/*.so; do echo "== $v"; SPZB200_LIB=$v python scripts/ply_sweep.py 2>&1 | grep '40000000'; done
