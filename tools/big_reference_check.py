#!/usr/bin/env python3
"""Scale check of the reference-side limits on a GPU box: a random reference of (default) 2.2 Gbp in two contigs - forward + reverse
size 4.4 G > 2^32, so global positions are 40 bits wide, and ~2.4 G index entries > 2^31, so the device index build runs in chunks of
block lengths.  Reads with known origin (1 % substitutions, both strands) must align to where they came from.
usage: XM_TRACE_SETUP=1 python tools/big_reference_check.py [total_bases] [n_reads]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mapper_b200 import capi, synth

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_200_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
L = 150
rng = np.random.default_rng(7)
lens = [total - total // 2 - 100_000_000, total // 2 + 100_000_000]   # the reference sorts contigs by length: shorter first
t0 = time.time()
contigs = []
for i, n in enumerate(lens):
    contigs.append(synth.CODES[rng.integers(0, 4, size=n, dtype=np.uint8)])
print("reference: %d + %d bases (%.1f s to generate)" % (lens[0], lens[1], time.time() - t0), flush=True)
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
t0 = time.time()
g.set_reference([synth.pack_contig(c) for c in contigs], lens)
print("set_reference: %.1f s" % (time.time() - t0), flush=True)
t0 = time.time(); g.build_index(L); print("index build (device): %.1f s" % (time.time() - t0), flush=True)
t0 = time.time(); g.build_duplications(-1, -1, 2, 1000); print("duplication table: %.1f s" % (time.time() - t0), flush=True)
mi, mb = g.index_info()
n_pos = sum(g.index_length_size(n)[1] for n in range(1, mb + 1))
print("index: lengths %d..%d, %d positions" % (mi, mb, n_pos), flush=True)
# reads with known origin
cidx = rng.integers(0, 2, size=n_reads)
pos = np.array([rng.integers(0, lens[c] - L) for c in cidx], dtype=np.int64)
strand = rng.integers(0, 2, size=n_reads)
reads = []
for i in range(n_reads):
    r = contigs[cidx[i]][pos[i]:pos[i] + L].copy()
    if strand[i]:
        r = synth.COMP[r][::-1]
    k = rng.binomial(L, 0.01)
    if k:
        w = rng.integers(0, L, size=k)
        r[w] = synth.CODES[(np.searchsorted(synth.CODES, r[w]) + rng.integers(1, 4, size=k)) % 4]
    reads.append(r)
batch = synth.batch_from_reads(reads)
t0 = time.time(); res = g.align_batch(batch, strict=True); dt = time.time() - t0
ok = 0; multi = 0; none = 0
assert (res["q_status"] == 0).all()
qc = res["q_comp_off"]; cc = res["comp_choice_off"]; cs = res["choice_sa_off"]; sb = res["sa_block_off"]; blocks = res["blocks"].reshape(-1, 4)
for q in range(n_reads):
    if qc[q + 1] == qc[q] or cc[qc[q] + 1] == cc[qc[q]]:
        none += 1
        continue
    n_choice = cc[qc[q] + 1] - cc[qc[q]]
    hit = False
    for ch in range(cc[qc[q]], cc[qc[q] + 1]):
        sa = cs[ch]
        b = blocks[sb[sa]]   # (start in the query, start in the contig, lengths)
        if res["sa_contig"][sa] == cidx[q] and bool(res["sa_reversed"][sa]) == bool(strand[q]) and abs(int(b[1]) - int(b[0]) - int(pos[q])) <= 3:
            hit = True
    ok += hit
    multi += n_choice > 1
print("aligned %d reads in %.2f s: %d at their origin (%.3f %%), %d unaligned, %d with several choices" % (n_reads, dt, ok, 100.0 * ok / n_reads, none, multi))
assert ok >= 0.995 * n_reads
print("OK")
