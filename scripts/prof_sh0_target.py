#!/usr/bin/env python
"""ncu target: two launches each of the SH-less kernels (config 4 of BASELINE.json): encode v3 with a LUF
flip, decode v3 to RUF, decode of a v2 (first-three quaternion) stream to LUF.  40M points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud
n, deg, dev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000, 0, torch.device("cuda", 0)
with codec.Context(0) as ctx:
    cloud = torch_cloud(n, deg, dev, seed=1)
    packed = codec.alloc_packed(n, deg, 3, device=dev)
    out = codec.alloc_cloud(n, deg, device=dev)
    rot = torch.randint(0, 256, (3 * n,), dtype=torch.uint8, device=dev)
    p2 = codec.PackedPlanes(n, deg, packed.positions, packed.scales, rot, packed.alphas, packed.colors, packed.sh, fractional_bits=12, version=2)
    for _ in range(2):
        ctx.encode_device(cloud, 7, out=packed)   # CoordinateSystem ids, splat-types.h:24-34: 7 = LUF, 8 = RUF
        ctx.decode_device(packed, 8, out=out)
        ctx.decode_device(p2, 7, out=out)
    torch.cuda.synchronize()
print("ok")
