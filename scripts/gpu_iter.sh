#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 120 scripts/_build/membench 100000000 | tee gpurun_out/membench.txt
timeout 600 python -m pytest tests/test_cxx_api.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2>gpurun_out/bench_tmp.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_tmp.json') if l.startswith('{')][-1])
print(d['value'], d['latency_60k'])"
