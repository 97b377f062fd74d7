#!/bin/bash
# first GPU contact: box facts, smoke, parity tests, small + full bench
mkdir -p gpurun_out
{ nproc; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null; nvidia-smi -L; lscpu | head -20; } > gpurun_out/box.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --points 10000000 --steps 20 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; echo "bench10m rc=$?"
cat gpurun_out/bench_10m.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_100m.json 2> gpurun_out/bench_100m.err; echo "bench100m rc=$?"
cat gpurun_out/bench_100m.json
tail -3 gpurun_out/bench_100m.err
