"""Where the first call of a process goes.  Fresh process, no torch: loads the
library, creates contexts with SPZB200_TRACE_INIT=1 (the library prints its own breakdown to stderr),
then times first / later calls through the pooled path a C++ API caller takes."""
import ctypes as C
import os
import sys
import time

t_proc = time.perf_counter()
os.environ["SPZB200_TRACE_INIT"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from spz_b200 import _native as N  # noqa: E402
from spz_b200 import codec  # noqa: E402

t0 = time.perf_counter()
L = N.lib()
print(f"dlopen libspz_b200.so                   {1e3 * (time.perf_counter() - t0):8.2f} ms")


def timed(label, fn):
    t = time.perf_counter()
    r = fn()
    print(f"{label:<40s}{1e3 * (time.perf_counter() - t):8.2f} ms", flush=True)
    return r


c1 = timed("context #1 (includes CUDA initialisation)", lambda: codec.Context(0))
c2 = timed("context #2 (spzb200_create, CUDA is up)", lambda: codec.Context(0))
c3 = timed("context #3", lambda: codec.Context(0))
p1 = timed("pooled lease #1 (spzb200_acquire, creates)", lambda: codec.Context(0, pooled=True))
p1.close()
p2 = timed("pooled lease #2 (reuses)", lambda: codec.Context(0, pooled=True))

from spz_b200.synth import numpy_cloud  # noqa: E402

for n in (60_000, 1_000_000):
    src = numpy_cloud(n, 3, seed=3)
    for rep in range(3):
        timed(f"encode_host {n} SH3 pageable, call {rep}", lambda: p2.encode_host(src, 6))
p2.close()
print(f"whole script                            {1e3 * (time.perf_counter() - t_proc):8.2f} ms")
