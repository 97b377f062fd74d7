"""Device-resident batched per-gaussian access (spzb200_unpack_gather_device): in-order, shuffled and strided index lists (development tool)."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
dev = torch.device("cuda", 0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
with codec.Context(0) as ctx:
    for deg in (3, 0):
        w = codec.byte_plane_widths(deg, 3)
        planes = [torch.randint(0, 256, (n * k,), dtype=torch.uint8, device=dev) for k in w]
        planes[2].view(torch.int32).bitwise_and_(-268698881)  # 0xEFFBFEFF: valid smallest-three payloads
        p = codec.PackedPlanes(n, deg, *planes, fractional_bits=12, version=3)
        out = torch.empty((n, 59), dtype=torch.float32, device=dev)
        for label, idx in (("in order (no index list)", None), ("shuffled", torch.randperm(n, device=dev)), ("every 7th, wrapped", (torch.arange(n, device=dev) * 7) % n)):
            for _ in range(2):
                ctx.unpack_gather_device(p, idx, None, out=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                e[0].record()
                ctx.unpack_gather_device(p, idx, None, out=out)
                e[1].record()
                torch.cuda.synchronize()
                ts.append(e[0].elapsed_time(e[1]))
            ms = statistics.median(ts)
            per = sum(w) + 236 + (8 if idx is not None else 0)
            print(json.dumps({"gaussians": n, "sh_degree": deg, "indices": label, "ms": round(ms, 3), "mgaussians_s": round(n / ms / 1e3), "hbm_gbs": round(per * n / ms / 1e6)}), flush=True)
        del planes, out
        torch.cuda.empty_cache()
