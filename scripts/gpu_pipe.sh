#!/bin/bash
export SPZB200_NO_REBUILD=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from spz_b200 import codec
from spz_b200.synth import numpy_cloud
n, deg = 10_000_000, 3
c = numpy_cloud(n, deg, 3)
with codec.Context(0) as ctx:
    for label, pinned, bounce, threads, chunk in (("pageable bounce t8 256K", False, True, 8, 1<<18), ("pageable bounce t4 256K", False, True, 4, 1<<18),
                                  ("pageable bounce t12 256K", False, True, 12, 1<<18), ("pageable bounce t8 128K", False, True, 8, 1<<17), ("pageable bounce t8 512K", False, True, 8, 1<<19),
                                  ("pageable unstaged", False, False, 0, 0), ("pinned", True, True, 0, 0)):
        ctx.set_host_staging(bounce, threads); ctx.set_chunk_points(chunk)
        src = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pinned)
        for a, b in zip(src.planes(), c.planes()): a[...] = b
        out = codec.alloc_packed(n, deg, 3, numpy_arrays=True, pinned=pinned)
        back = codec.alloc_cloud(n, deg, numpy_arrays=True, pinned=pinned)
        for a in list(out.planes()) + list(back.planes()): a[...] = 0
        best = [1e9, 1e9]
        for rep in range(4):
            t0 = time.perf_counter(); _, te = ctx.encode_host(src, 6, out=out); t1 = time.perf_counter()
            _, td = ctx.decode_host(out, 6, out=back); t2 = time.perf_counter()
            if rep == 0: first = (t1 - t0, t2 - t1)
            best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
        print("%-26s encode %.1f ms (%.1f GB/s) decode %.1f ms (%.1f GB/s)  first call %.0f/%.0f ms  host_copy %.0f/%.0f ms" % (label, best[0]*1e3, n*236/best[0]/1e9, best[1]*1e3, n*236/best[1]/1e9, first[0]*1e3, first[1]*1e3, te["host_copy_ms"], td["host_copy_ms"]), flush=True)
PY
