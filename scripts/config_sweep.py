#!/usr/bin/env python
"""Device-timed kernels of the BASELINE.json configs that bench.py does not headline (development
tool; bench.py measures configs[4], the 100M SH3 cloud): 10M SH3 v3 (config 3) and 10M SH0 with
LUF / RUF conversion, v3 and the v2 first-three decode (config 4).  One JSON line per case."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud

def timed(fn, reps=20, rounds=5):
    """Median over `rounds` of (time of `reps` launches queued back to back) / reps: the stream never
    runs dry, so the host's launch path (tens of microseconds through Python) stays outside the
    figure; every working set here exceeds the 126 MB L2."""
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

LUF, RUF = 7, 8  # splat-types.h:24-34 (round 1 passed 3 and 4 here, which are LUB and RUB -- RUB is "no flip at all")
dev = torch.device("cuda", 0)
with codec.Context(0) as ctx:
    for label, n, deg, ver, frames in (("config 3: 10M SH3 v3", 10_000_000, 3, 3, (6,)), ("config 4: 10M SH0 v3", 10_000_000, 0, 3, (LUF, RUF)),
                                       ("config 4: 10M SH0 v2 decode", 10_000_000, 0, 2, (LUF, RUF)), ("10M SH1 v3", 10_000_000, 1, 3, (6,)),
                                       ("10M SH2 v3", 10_000_000, 2, 3, (6,)), ("2.5M SH3 v3", 2_500_000, 3, 3, (6,)),
                                       ("100M SH0 v3", 100_000_000, 0, 3, (LUF,)), ("100M SH0 v2 decode", 100_000_000, 0, 2, (RUF,))):
        cloud = torch_cloud(n, deg, dev, seed=1)
        packed = codec.alloc_packed(n, deg, 3, device=dev)
        out = codec.alloc_cloud(n, deg, device=dev)
        for frame in frames:  # CoordinateSystem ids (splat-types.h:24-34): 6 = RDF, 7 = LUF, 8 = RUF
            res = {"case": label, "points": n, "sh_degree": deg, "stream_version": ver, "coordinate_system": frame,
                   "coordinate_name": {6: "RDF", 7: "LUF", 8: "RUF"}[frame], "flip_bits_encode": codec.flip_bits(frame, 4), "flip_bits_decode": codec.flip_bits(4, frame)}
            if ver == 3:
                b = codec.algorithmic_bytes_per_gaussian(deg, 3) * n
                ms = timed(lambda: ctx.encode_device(cloud, frame, out=packed))
                res["encode"] = {"ms": round(ms, 4), "hbm_gbs": round(b / ms / 1e6), "mgaussians_s": round(n / ms / 1e3)}
                ms = timed(lambda: ctx.decode_device(packed, frame, out=out))
                res["decode"] = {"ms": round(ms, 4), "hbm_gbs": round(b / ms / 1e6), "mgaussians_s": round(n / ms / 1e3)}
            else:
                # any bytes are a valid v2 stream (load-spz.cc:344 clamps): v3 planes with a 3-byte rotation plane
                ctx.encode_device(cloud, frame, out=packed)
                rot = torch.randint(0, 256, (3 * n,), dtype=torch.uint8, device=dev)
                p2 = codec.PackedPlanes(n, deg, packed.positions, packed.scales, rot, packed.alphas, packed.colors, packed.sh,
                                        fractional_bits=12, version=2)
                b = codec.algorithmic_bytes_per_gaussian(deg, 2) * n
                ms = timed(lambda: ctx.decode_device(p2, frame, out=out))
                res["decode"] = {"ms": round(ms, 4), "hbm_gbs": round(b / ms / 1e6), "mgaussians_s": round(n / ms / 1e3)}
            print(json.dumps(res), flush=True)
        del cloud, packed, out
        torch.cuda.empty_cache()
