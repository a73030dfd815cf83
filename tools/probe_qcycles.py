#!/usr/bin/env python3
"""Profiling aid: per-query SM-clock cost distribution of the align kernel (XM_QCYCLES=1) on the bench workload.
Usage (GPU box): python tools/probe_qcycles.py [--reads N] [--paired] > gpurun_out/qcycles.txt"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["XM_QCYCLES"] = "1"
from mapper_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--ref-bases", type=int, default=5000000)
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--paired", action="store_true")
a = ap.parse_args()
ref = synth.random_reference(a.ref_bases, seed=1)
batch = synth.simulate_reads_fast(ref, a.reads, a.read_len, seed=2, paired=a.paired, inner_mean=300.0, inner_sd=30.0, per_penalty=50.0)
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
g.set_reference([synth.pack_contig(s) for _, s in ref], [len(s) for _, s in ref])
g.build_index(a.read_len)
g.build_duplications(-1, -1, 2, 1000)
for it in range(2):
    r = g.align_batch(batch)
st = r["stats"]
S = capi.STAT
print("easy pass queries:", int(st[S["easy"]]), "ms:", st[S["easy_ns"]] / 1e6)
print("tiers queries:", [int(st[S["tier%d" % t]]) for t in range(3)], "tier ms:", [st[S["tier%d_ns" % t]] / 1e6 for t in range(3)], "total kernel ms:", st[S["kernel_ns"]] / 1e6)
print("probes %d seeds %d hits %d straight %d path_calls %d path_steps %d path_cells %d" % tuple(int(st[S[k]]) for k in ["probes", "seeds", "hits", "straight", "path_calls", "path_steps", "path_cells"]))
print("phase cycles (sum over queries): " + "  ".join("%s=%.3e" % (k, float(st[S["cyc_" + k]])) for k in ["seed", "straight", "hba", "path", "tables", "spare", "total"]))
cy = r["q_cycles"].astype(np.float64)
nq = len(cy)
# classify reads by result shape
first_choice = r["comp_choice_off"][r["q_comp_off"][:-1]]
n_choice = r["comp_choice_off"][r["q_comp_off"][:-1] + 1] - first_choice
blocks_per_sa = np.diff(r["sa_block_off"])
sa_first = r["choice_sa_off"][np.minimum(first_choice, len(r["choice_sa_off"]) - 2)]
nblk = np.where(n_choice > 0, blocks_per_sa[np.minimum(sa_first, len(blocks_per_sa) - 1)], 0)
pen = np.where(n_choice > 0, r["choice_f64"].reshape(-1, 4)[np.minimum(first_choice, len(r["choice_inner"]) - 1), 3], -1)


def pct(x, name):
    if len(x) == 0:
        print("%-34s n=0" % name)
        return
    q = np.percentile(x, [50, 90, 99, 99.9, 100])
    print("%-34s n=%7d  sum=%.3e  mean=%9.0f  p50=%9.0f p90=%9.0f p99=%9.0f p99.9=%10.0f max=%10.0f" % (name, len(x), x.sum(), x.mean(), *q))


pct(cy, "all (cycles of finishing tier)")
pct(cy[n_choice == 0], "unaligned")
pct(cy[(n_choice == 1) & (nblk == 1) & (pen == 0)], "1 choice, ungapped, penalty 0")
pct(cy[(n_choice == 1) & (nblk == 1) & (pen > 0) & (pen <= 2)], "1 choice, ungapped, 0<pen<=2")
pct(cy[(n_choice == 1) & (nblk == 1) & (pen > 2)], "1 choice, ungapped, pen>2")
pct(cy[(n_choice == 1) & (nblk > 1)], "1 choice, gapped")
pct(cy[n_choice > 1], ">1 choices")
top = np.argsort(-cy)[:10]
print("top-10 queries:", [(int(i), int(cy[i]), int(n_choice[i]), int(nblk[i]), float(pen[i])) for i in top])
g.close()
