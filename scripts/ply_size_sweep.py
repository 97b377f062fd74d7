#!/usr/bin/env python
"""As size_sweep.py, for the fused PLY-rows kernels (canonical property order)."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec

def timed(fn, reps=20, rounds=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for _ in range(reps):
            fn()
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]) / reps)
    return statistics.median(ts)

deg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "6e4,3e5,6e5,1.25e6,2.5e6,5e6,1e7".split(","))]
names = codec.ply_property_names(deg)
w = len(names)
with codec.Context(0) as ctx:
    for n in sizes:
        rows = torch.empty(n * w, dtype=torch.float32, device="cuda").uniform_(-1, 1)
        out = codec.alloc_packed(n, deg, 3, device="cuda")
        back = torch.empty_like(rows)
        b = (4 * w + codec.packed_bytes_per_gaussian(deg)) * n
        e = timed(lambda: ctx.encode_ply_device(rows, n, names, deg, 6, out=out))
        d = timed(lambda: ctx.decode_ply_device(out, names, 6, out=back))
        print(json.dumps({"points": n, "sh_degree": deg, "rows_to_packed_us": round(e * 1e3, 1), "packed_to_rows_us": round(d * 1e3, 1),
                          "rows_to_packed_gbs": round(b / e / 1e6), "packed_to_rows_gbs": round(b / d / 1e6)}), flush=True)
        del rows, out, back
        torch.cuda.empty_cache()
