#!/usr/bin/env python
"""Device-timed size / degree sweep of the two codec kernels (development tool, not the bench).
Prints one JSON line per (points, sh_degree): GB/s of encode and decode, per-step spread."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spz_b200 import codec
from spz_b200.synth import torch_cloud


def run(ctx, n, deg, steps=10, ver=3):
    dev = torch.device("cuda", 0)
    cloud = torch_cloud(n, deg, dev, seed=1)
    packed = codec.alloc_packed(n, deg, 3, device=dev)
    out = codec.alloc_cloud(n, deg, device=dev)
    for _ in range(3):
        ctx.encode_device(cloud, 6, out=packed)
        ctx.decode_device(packed, 6, out=out)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for s in range(steps):
        ev[s][0].record()
        ctx.encode_device(cloud, 6, out=packed)
        ev[s][1].record()
        ctx.decode_device(packed, 6, out=out)
        ev[s][2].record()
    torch.cuda.synchronize()
    enc = [e[0].elapsed_time(e[1]) for e in ev]
    dec = [e[1].elapsed_time(e[2]) for e in ev]
    b = codec.algorithmic_bytes_per_gaussian(deg, 3) * n
    g = lambda ms: b / (ms * 1e-3) / 1e9  # noqa: E731
    print(json.dumps({"points": n, "sh_degree": deg, "enc_gbs": round(g(statistics.median(enc))), "dec_gbs": round(g(statistics.median(dec))),
                      "enc_ms": [round(x, 3) for x in enc[:6]], "dec_ms": [round(x, 3) for x in dec[:6]],
                      "enc_best_gbs": round(g(min(enc))), "dec_best_gbs": round(g(min(dec)))}), flush=True)
    del cloud, packed, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2.5e6", "1e7", "4e7", "1e8"])]
    degs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["3"])]
    with codec.Context(0) as ctx:
        for deg in degs:
            for n in sizes:
                run(ctx, n, deg)
