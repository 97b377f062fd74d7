"""A/B of launch-level knobs over cloud sizes of real scenes: every combination of the given environment
settings runs scripts/size_sweep.py in a process of its own (the knobs are read at context creation).

    python scripts/ab_sweep.py "SPZB200_PDL=0,1;SPZB200_REST=separate,fold" 1.25e6,2.5e6,5e6,1e7 0,1,3
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
knobs = [(k.split("=")[0], k.split("=")[1].split(",")) for k in sys.argv[1].split(";") if k]
sizes = sys.argv[2] if len(sys.argv) > 2 else "1.25e6,2.5e6,5e6,1e7"
degs = [int(d) for d in (sys.argv[3] if len(sys.argv) > 3 else "0,1,3").split(",")]
for combo in itertools.product(*[v for _, v in knobs]):
    setting = {k: v for (k, _), v in zip(knobs, combo)}
    env = dict(os.environ, SPZB200_NO_REBUILD="1")
    for k, v in setting.items():
        if v == "-":
            env.pop(k, None)
        else:
            env[k] = v
    for deg in degs:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "size_sweep.py"), sizes, str(deg)], capture_output=True, text=True, env=env)
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                print(json.dumps({**setting, **json.loads(ln)}), flush=True)
        if r.returncode != 0:
            print(json.dumps({**setting, "sh_degree": deg, "error": r.stderr[-400:]}), flush=True)
