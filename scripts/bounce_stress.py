#!/usr/bin/env python
"""Development tool: randomized stress of the host-pointer pipeline.  Two host threads, one context each, run
encode_host / decode_host on clouds of random size, SH degree, copy-thread count, range size and pinned-ness for
`seconds`; every result is compared, plane by plane, with what the device-resident entry points produce for the
same cloud (those are pinned to the oracle by the tests).  Prints one JSON line; "mismatches" must be 0."""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda", 0)
stats = {"calls": 0, "mismatches": 0, "gaussians": 0, "errors": []}
lock = threading.Lock()

def worker(seed):
    rng = np.random.default_rng(seed)
    with codec.Context(0) as ctx, codec.Context(0) as ref:
        t_end = time.time() + seconds
        while time.time() < t_end:
            deg = int(rng.integers(0, 4))
            n = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 400_000), rng.integers(400_000, 3_000_000)]))
            threads = int(rng.choice([0, 1, 2, 5, 12]))
            chunk = int(rng.choice([0, 0, 65536, 131072, 1 << 20]))
            pin_in, pin_out = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            try:
                ctx.set_host_staging(int(rng.choice([1, 2])), threads)
                os.environ.pop("SPZB200_PAGEABLE_CHUNK_POINTS", None)
                cloud = torch_cloud(n, deg, dev, seed=int(rng.integers(1, 1 << 30)))
                want_p = ref.encode_device(cloud, 7)
                want_c = ref.decode_device(want_p, 5)
                torch.cuda.synchronize()
                src = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pin_in)
                for a, b in zip(src.planes(), cloud.planes()):
                    a[...] = b.cpu().numpy()
                out = codec.alloc_packed(n, deg, 3, numpy_arrays=True, pinned=pin_out)
                if chunk:
                    ctx.set_chunk_points(chunk)
                got_p, _ = ctx.encode_host(src, 7, out=out)
                pk = codec.alloc_packed(n, deg, 3, numpy_arrays=True, pinned=pin_in)
                for a, b in zip(pk.planes(), want_p.planes()):
                    a[...] = b.cpu().numpy()
                back = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pin_out)
                got_c, _ = ctx.decode_host(pk, 5, out=back)
                ctx.set_chunk_points(0)
                bad = 0
                for a, b in zip(got_p.planes(), want_p.planes()):
                    bad += int(not np.array_equal(np.asarray(a), b.cpu().numpy()))
                for a, b in zip(got_c.planes(), want_c.planes()):
                    bad += int(not np.array_equal(np.asarray(a).view(np.uint32), b.cpu().numpy().view(np.uint32)))
                with lock:
                    stats["calls"] += 2
                    stats["gaussians"] += 2 * n
                    stats["mismatches"] += bad
                    if bad:
                        stats["errors"].append({"n": n, "deg": deg, "threads": threads, "chunk": chunk, "pinned": [pin_in, pin_out]})
            except Exception as e:  # noqa: BLE001 -- a stress tool reports, it does not stop
                with lock:
                    stats["errors"].append({"exception": repr(e)[:300], "n": n, "deg": deg})
                    stats["mismatches"] += 1

ts = [threading.Thread(target=worker, args=(s,)) for s in (11, 22)]
[t.start() for t in ts]
[t.join() for t in ts]
stats["errors"] = stats["errors"][:10]
print(json.dumps(stats))
