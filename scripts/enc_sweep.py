"""Encode-only timing at a few sizes (development): python scripts/enc_sweep.py <sizes> <degrees>; honours the SPZB200_* knobs."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud
dev = torch.device("cuda", 0)
with codec.Context(0) as ctx:
    for deg in [int(d) for d in sys.argv[2].split(",")]:
        for n in [int(float(x)) for x in sys.argv[1].split(",")]:
            cloud = torch_cloud(n, deg, dev, seed=1)
            packed = codec.alloc_packed(n, deg, 3, device=dev)
            for _ in range(3):
                ctx.encode_device(cloud, 6, out=packed)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                e[0].record()
                for _ in range(10):
                    ctx.encode_device(cloud, 6, out=packed)
                e[1].record()
                torch.cuda.synchronize()
                ts.append(e[0].elapsed_time(e[1]) / 10)
            ms = statistics.median(ts)
            print(json.dumps({"points": n, "sh_degree": deg, "encode_us": round(ms * 1e3, 1), "encode_gbs": round(codec.algorithmic_bytes_per_gaussian(deg, 3) * n / ms / 1e6)}), flush=True)
            del cloud, packed
            torch.cuda.empty_cache()
