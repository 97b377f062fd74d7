#!/bin/bash
export SPZB200_NO_REBUILD=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import torch, time
n = 2_400_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
print("H2D %.1f GB/s" % (n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H %.1f GB/s" % (n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidirectional %.1f GB/s each way" % (n / t(both) / 1e9))
PY
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1d.json') if l.startswith('{')][-1])
print(d['value'], d['roofline']['encode']['achieved'], d['roofline']['decode']['achieved'], d['e2e']['value'], json.dumps(d.get('host_zlib')))"
