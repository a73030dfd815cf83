#!/usr/bin/env python3
"""Static SASS statistics of one device function inside a kernel image of a shared library (no GPU needed).
usage: python tools/sass_fn_stats.py lib.so kernel_substr function_substr [--dump]"""
import collections, os, re, subprocess, sys, tempfile
lib, ksub, fsub = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
sym = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
rng = None
for ln in sym.splitlines():
    p = ln.split()
    if len(p) >= 8 and p[3] == "FUNC" and ksub in p[7] and fsub in p[7] and "$" in p[7]:
        rng = (int(p[1], 16), int(p[2])); name = p[7]
if rng is None: sys.exit("function not found")
dis = subprocess.run(["cuobjdump", "-sass", os.path.abspath(lib)], capture_output=True, text=True).stdout
inside = False; ops = collections.Counter(); n = 0; lines = []
for ln in dis.splitlines():
    if "Function :" in ln: inside = ksub in ln and "$" not in ln
    if not inside: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m: continue
    a = int(m.group(1), 16)
    if rng[0] <= a < rng[0] + rng[1]:
        txt = m.group(2).strip(); t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1; n += 1; lines.append("%6x  %s" % (a, txt))
print(name, "bytes", rng[1], "instructions", n)
print(dict(ops.most_common(16)))
if "--dump" in sys.argv: print("\n".join(lines))
