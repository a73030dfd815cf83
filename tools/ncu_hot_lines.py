#!/usr/bin/env python3
"""Maps the per-SASS-instruction samples of an .ncu-rep back to source lines with nvdisasm line info of the SAME library.
usage: python tools/ncu_hot_lines.py report.ncu-rep libxmapper_b200.so [launch_index] [top_n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# per kernel section: list of (offset, file, line, inline-chain function)
sections = {}
cur_sec = None; cur = None
for ln in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        cur_sec = m.group(1); sections[cur_sec] = {}; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m and cur_sec is not None:
        sections[cur_sec][int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')[1:]
blk = blocks[launch]
rows = list(csv.reader(('"Kernel Name"' + blk).splitlines()))
kname = rows[0][1]
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
# find the matching section: mangled name contains "xm_align_kernelILb1" for <1>
want = "ILb1" if ("<1>" in kname or "(bool)1" in kname) else "ILb0" if ("<0>" in kname or "(bool)0" in kname) else ""
sec = [k for k in sections if want in k and "align" in k]
sec = sections[sec[0]] if sec else max(sections.values(), key=len)
base = min(int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]]) for r in data)
samp = collections.Counter(); inst = collections.Counter(); noinst = collections.Counter()
tot = 0; toti = 0
for r in data:
    a = r[col["Address"]]
    a = (int(a, 16) if a.startswith("0x") else int(a)) - base
    key = sec.get(a, ("?", 0))
    s = float(r[col["# Samples"]] or 0); i = float(r[col["Instructions Executed"]] or 0)
    samp[key] += s; inst[key] += i; noinst[key] += float(r[col["stall_no_inst"]] or 0)
    tot += s; toti += i
print("# %s  launch %d: %s   total samples %.0f, warp instructions %.4e" % (rep, launch, kname, tot, toti))
print("# by samples")
for k, v in samp.most_common(topn):
    print("%6.2f%% samples  %6.2f%% inst  no_inst %5.1f%%  %s:%d" % (100 * v / tot, 100 * inst[k] / toti, 100 * noinst[k] / max(v, 1), k[0], k[1]))
print("# by instructions executed")
for k, v in inst.most_common(topn):
    print("%6.2f%% inst  %s:%d" % (100 * v / toti, k[0], k[1]))
