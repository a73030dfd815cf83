#!/usr/bin/env python3
"""Builds the hash-block index on the device for a synthetic reference (GPU box): python tools/probe_index.py [--ref-bases N] [--contigs C]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapper_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-bases", type=int, default=50000000)
ap.add_argument("--contigs", type=int, default=10)
ap.add_argument("--max-used", type=int, default=150)
a = ap.parse_args()
ref = synth.random_reference(a.ref_bases, seed=4, n_contigs=a.contigs, repeat_fraction=0.05, repeat_copies=(2, 4), repeat_len=(1000, 5000))
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
g.set_reference([synth.pack_contig(s) for _, s in ref], [len(s) for _, s in ref])
for it in range(2):
    t0 = time.time()
    g.build_index(a.max_used)
    print("device index build, %d bases, lengths up to %d: %.3f s" % (a.ref_bases, g.index_info()[1], time.time() - t0), flush=True)
n_pos = sum(len(g.get_index_length(n)["positions"]) for n in range(1, g.index_info()[1] + 1))
print("positions stored: %d" % n_pos)
