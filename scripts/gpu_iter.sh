#!/bin/bash
export SPZB200_NO_REBUILD=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cxx_api.py -m gpu -x -q -k "ply" 2>&1 | tail -3
python scripts/ply_sweep.py 2>&1 | grep "packed->rows"
