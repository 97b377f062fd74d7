"""cuobjdump -sass of one kernel of the shipped library with an opcode histogram on top (what profiles/r*_sass_*.txt hold).
usage: sass_profile.py <mangled-name substring> "<description>" > profiles/rN_sass_X.txt     (no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "spz_b200", "_lib", "libspz_b200.so")
pat, desc = sys.argv[1], sys.argv[2]
names = [ln.split("Function : ")[1].strip() for ln in subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
         if "Function : " in ln and pat in ln]
assert names, f"no kernel matches {pat}"
name = names[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
body = sass[sass.index("Function :"):]
kernel_only = body.split(".text.", 1)[0]
ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body)
full, base = collections.Counter(ops), collections.Counter(o.split(".")[0] for o in ops)
special = {k: v for k, v in full.items() if re.match(r"LDG|STG|LDS|STS|UBLKCP|SYNCS|MUFU|PRMT|SHF|I2IP|F2IP|ACQBULK|LDL|STL|CALL|BAR|ATOM|RED|CCTL|MEMBAR|ERRBAR|LDC|ELECT|PREEXIT|ACQ", k)}
print(f"# cuobjdump -sass of {desc} ({name}) in spz_b200/_lib/libspz_b200.so, sm_100a")
print(f"# {len(ops)} instructions; opcode histogram: " + ", ".join(f"{k} {v}" for k, v in base.most_common(24)))
print("# memory / special opcodes: " + ", ".join(f"{k} {v}" for k, v in sorted(special.items())))
print(body)
