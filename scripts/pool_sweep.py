"""Short-lived host threads against the context pool (spzb200_acquire): 64 threads that each pack one cloud
and exit, in turn and all at once, for several pool sizes (SPZ_B200_MAX_CONTEXTS is read when the pool is
first used, so every setting runs in a process of its own).  Development tool; uses the test shim."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
import numpy as np
from test_cxx_api import Shim
from util import random_cloud
mine = Shim("b200")
out = {"max_contexts": os.environ.get("SPZ_B200_MAX_CONTEXTS", "default")}
for n in (60_000, 400_000):
    c = random_cloud(np.random.default_rng(1), n, 3, False)
    mine.short_lived_threads(c, 6, 8, True)   # warm: contexts and their buffers exist
    mine.short_lived_threads(c, 6, 8, True)
    out[str(n)] = {"one_thread_64_packs_ms": min(mine.short_lived_threads(c, 6, 64, 2) for _ in range(2)),
                   "64_threads_in_turn_ms": mine.short_lived_threads(c, 6, 64, False),
                   "64_threads_at_once_ms": min(mine.short_lived_threads(c, 6, 64, True) for _ in range(2))}
print(json.dumps(out))
"""
for k in ("1", "2", "4", "8"):
    env = dict(os.environ, REPO=ROOT, SPZ_B200_MAX_CONTEXTS=k, SPZB200_NO_REBUILD="1")
    r = subprocess.run([sys.executable, "-c", WORKER], capture_output=True, text=True, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    print(lines[-1] if lines else json.dumps({"max_contexts": k, "error": r.stderr[-500:]}), flush=True)
