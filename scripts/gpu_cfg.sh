#!/bin/bash
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
timeout 2400 python -m pytest tests/test_baseline_configs.py -m gpu -x -q --durations=10 > gpurun_out/pytest_cfg.log 2>&1; echo "cfg rc=$?"; tail -25 gpurun_out/pytest_cfg.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1c.json') if l.startswith('{')][-1])
print(d['value'], d['roofline']['encode'], d['roofline']['decode'], d['e2e']['value'], d.get('host_zlib'))"
