#!/bin/bash
export SPZB200_NO_REBUILD=1
echo "== default"; python scripts/kernel_sweep.py 1e8 3,0 2>&1 | cut -c1-95
for v in spz_b200/_lib/variants/*.so; do echo "== $v"; SPZB200_LIB=$v python scripts/kernel_sweep.py 1e8 3,0 2>&1 | cut -c1-95; done
echo "== pageable vs pinned host pipeline, 10M SH3"
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from spz_b200 import codec
from spz_b200.synth import numpy_cloud
n, deg = 10_000_000, 3
c = numpy_cloud(n, deg, 3)
with codec.Context(0) as ctx:
    for label, pinned in (("pageable", False), ("pinned", True)):
        src = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pinned)
        for a, b in zip(src.planes(), c.planes()): a[...] = b
        out = codec.alloc_packed(n, deg, 3, numpy_arrays=True, pinned=pinned)
        back = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pinned)
        for a in list(out.planes()) + list(back.planes()): a[...] = 0   # fault the pages in
        for rep in range(3):
            t0 = time.perf_counter(); _, te = ctx.encode_host(src, 6, out=out); t1 = time.perf_counter()
            _, td = ctx.decode_host(out, 6, out=back); t2 = time.perf_counter()
        print(label, "encode %.1f ms (%.1f GB/s in)" % ((t1 - t0) * 1e3, n * 236 / (t1 - t0) / 1e9), "decode %.1f ms (%.1f GB/s out)" % ((t2 - t1) * 1e3, n * 236 / (t2 - t1) / 1e9), flush=True)
PY
