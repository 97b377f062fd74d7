"""Development tool: full-duplex host pipeline (encode_host || decode_host on two contexts) vs range size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud
n, deg = 40_000_000, 3
c = torch_cloud(n, deg, "cuda", seed=3)
hc = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
for s, d in zip(c.planes(), hc.planes()): torch.from_numpy(d).copy_(s)
del c; torch.cuda.empty_cache()
hp = [codec.alloc_packed(n, deg, 3, pinned=True, numpy_arrays=True) for _ in range(2)]
hb = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
pool = ThreadPoolExecutor(2)
with codec.Context(0) as a, codec.Context(0) as b:
    a.encode_host(hc, 6, out=hp[1])
    for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
        a.set_chunk_points(chunk); b.set_chunk_points(chunk)
        def step(i):
            fa = pool.submit(a.encode_host, hc, 6, hp[i % 2]); fb = pool.submit(b.decode_host, hp[(i + 1) % 2], 6, hb)
            fa.result(); fb.result()
        step(0)
        t0 = time.perf_counter()
        for i in range(3): step(1 + i)
        dt = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for i in range(2):
            a.encode_host(hc, 6, out=hp[0]); a.decode_host(hp[0], 6, out=hb)
        ds = (time.perf_counter() - t0) / 2
        print(f"range {chunk:>8} points: duplex {dt*1e3:7.1f} ms ({n*301/dt/1e9:5.1f} GB/s each way)   sequential {ds*1e3:7.1f} ms", flush=True)
