#!/usr/bin/env python
"""Device-timed encode / decode over cloud sizes typical of real scenes (development tool).  Launches are
queued back to back (the stream never runs dry), `reps` alternating encode/decode pairs per measurement;
below ~400K gaussians the planes fit the 126 MB L2, so those figures are L2-, not HBM-bound."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spz_b200 import codec
from spz_b200.synth import torch_cloud

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

def timed_graph(fn, reps=20, rounds=5):
    """The same `reps` launches captured once into a CUDA graph and replayed: no host launch path at all between the
    kernels (the Python -> ctypes -> cudaLaunchKernelEx path costs ~10-20 us per call, which is longer than a kernel on
    a cloud of a million SH-less gaussians, so `timed` reports the host's launch rate there, not the kernel)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        g.replay()
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]) / reps)
    return statistics.median(ts)

use_graph = "--graph" in sys.argv
if use_graph:
    sys.argv.remove("--graph")
    timed = timed_graph
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "6e4,3e5,6e5,1.25e6,2.5e6,5e6,1e7,4e7".split(","))]
dev = torch.device("cuda", 0)
with codec.Context(0) as ctx:
    for n in sizes:
        cloud = torch_cloud(n, deg, dev, seed=1)
        packed = codec.alloc_packed(n, deg, 3, device=dev)
        out = codec.alloc_cloud(n, deg, device=dev)
        b = codec.algorithmic_bytes_per_gaussian(deg, 3) * n
        e = timed(lambda: ctx.encode_device(cloud, 6, out=packed))
        d = timed(lambda: ctx.decode_device(packed, 6, out=out))
        print(json.dumps({"timing": "cuda graph of 20 launches" if use_graph else "20 launches queued from Python", "points": n, "sh_degree": deg, "encode_us": round(e * 1e3, 1), "decode_us": round(d * 1e3, 1),
                          "encode_gbs": round(b / e / 1e6), "decode_gbs": round(b / d / 1e6)}), flush=True)
        del cloud, packed, out
        torch.cuda.empty_cache()
