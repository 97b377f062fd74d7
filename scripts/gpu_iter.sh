#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2>gpurun_out/bench_tmp.err; echo rc=$?; tail -3 gpurun_out/bench_tmp.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_tmp.json') if l.startswith('{')][-1])
print(d['value']); print(json.dumps(d['e2e'], indent=1)[:1500])"
