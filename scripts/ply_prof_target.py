#!/usr/bin/env python
"""One warm-up and one measured launch of each canonical-layout PLY kernel (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
n, deg = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 3
names = codec.ply_property_names(deg)
with codec.Context(0) as ctx:
    rows = torch.empty(n * len(names), dtype=torch.float32, device="cuda").uniform_(-1, 1)
    out = codec.alloc_packed(n, deg, 3, device="cuda")
    back = torch.empty_like(rows)
    for _ in range(2):
        ctx.encode_ply_device(rows, n, names, deg, 6, out=out)
        ctx.decode_ply_device(out, names, 6, out=back)
    torch.cuda.synchronize()
print("ok")
