#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/ceilings.txt 2>&1 <<'PY'
import torch
n = 6_000_000_000  # 24 GB of float32
a = torch.empty(n, dtype=torch.float32, device='cuda'); b = torch.empty(n, dtype=torch.float32, device='cuda')
def t(fn, reps=5):
    for _ in range(2): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(reps):
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize(); best = min(best, ev[0].elapsed_time(ev[1]))
    return best
print("# torch elementwise kernels on 24 GB float32 tensors, best of 5, CUDA events")
print("memset a.zero_() (write only)     %.0f GB/s" % (4 * n / t(lambda: a.zero_()) / 1e6))
print("fill a.fill_(1.5) (write only)    %.0f GB/s" % (4 * n / t(lambda: a.fill_(1.5)) / 1e6))
print("copy b.copy_(a) (read+write)      %.0f GB/s" % (8 * n / t(lambda: b.copy_(a)) / 1e6))
print("sum a.sum() (read only)           %.0f GB/s" % (4 * n / t(lambda: a.sum()) / 1e6))
print("mul_ a.mul_(2) (read+write)       %.0f GB/s" % (8 * n / t(lambda: a.mul_(2.0)) / 1e6))
PY
cat gpurun_out/ceilings.txt
