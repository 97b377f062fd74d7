"""Times scripts/size_sweep.py under each tuning-variant library in scripts/_build/variants (SPZB200_LIB) and the shipped one.
usage: variant_sweep.py <sizes> <degrees> [variant names...]"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes, degs = sys.argv[1], sys.argv[2].split(",")
names = sys.argv[3:] or ["shipped"] + sorted(os.path.basename(p)[7:-3] for p in glob.glob(os.path.join(ROOT, "scripts", "_build", "variants", "libspz_*.so")))
for rep in range(2):
    for name in names:
        env = dict(os.environ, SPZB200_NO_REBUILD="1")
        if name != "shipped":
            env["SPZB200_LIB"] = os.path.join(ROOT, "scripts", "_build", "variants", f"libspz_{name}.so")
        for deg in degs:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "size_sweep.py"), sizes, deg], capture_output=True, text=True, env=env)
            for ln in r.stdout.splitlines():
                if ln.startswith("{"):
                    print(json.dumps({"variant": name, **json.loads(ln)}), flush=True)
            if r.returncode:
                print(json.dumps({"variant": name, "error": r.stderr[-300:]}), flush=True)
