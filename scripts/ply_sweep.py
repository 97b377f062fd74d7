#!/usr/bin/env python
"""Device-timed fused PLY-rows encoder (development tool)."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec

def run(ctx, n, deg):
    names = codec.ply_property_names(deg)
    w = len(names)
    rows = torch.empty(n * w, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    out = codec.alloc_packed(n, deg, 3, device="cuda")
    for _ in range(3):
        ctx.encode_ply_device(rows, n, names, deg, 6, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); ctx.encode_ply_device(rows, n, names, deg, 6, out=out); e[1].record(); torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
    b = (4 * w + codec.packed_bytes_per_gaussian(deg)) * n
    ms = statistics.median(ts)
    print(json.dumps({"op": "rows->packed", "points": n, "sh_degree": deg, "bytes_per_gaussian": 4 * w + codec.packed_bytes_per_gaussian(deg), "ms": round(ms, 3), "gbs": round(b / ms / 1e6), "mgs": round(n / ms / 1e3)}), flush=True)
    back = torch.empty_like(rows)
    for _ in range(3):
        ctx.decode_ply_device(out, names, 6, out=back)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); ctx.decode_ply_device(out, names, 6, out=back); e[1].record(); torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
    ms = statistics.median(ts)
    print(json.dumps({"op": "packed->rows", "points": n, "sh_degree": deg, "ms": round(ms, 3), "gbs": round(b / ms / 1e6), "mgs": round(n / ms / 1e3)}), flush=True)

if __name__ == "__main__":
    sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1e7", "4e7"])]
    degs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["3", "0"])]
    with codec.Context(0) as ctx:
        for deg in degs:
            for n in sizes:
                run(ctx, n, deg)
