#!/bin/bash
# round-end style validation: GPU tests, smoke, both bench arms, C++ API timing
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_n1.json
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
bash scripts/gpu_api.sh 2>&1 | grep -v "rep\": [023]" | tee gpurun_out/api.log
