#!/usr/bin/env python3
"""BASELINE.json configs[3] shape at reduced scale (GPU box): multi-contig reference with repeat families, reads checked against
the CPU oracle bit for bit.  python tools/parity_large.py [--ref-bases N] [--contigs C] [--reads R]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import xm_oracle as xo  # noqa: E402
from mapper_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-bases", type=int, default=60000000)
ap.add_argument("--contigs", type=int, default=12)
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--paired", action="store_true")
ap.add_argument("--host-index", action="store_true", help="also time the host index builder")
a = ap.parse_args()
t0 = time.time()
ref = synth.random_reference(a.ref_bases, seed=4, n_contigs=a.contigs, repeat_fraction=0.05, repeat_copies=(2, 4), repeat_len=(1000, 5000))
print("reference: %d contigs, %d bases (%.1f s)" % (len(ref), sum(len(s) for _, s in ref), time.time() - t0), flush=True)
t0 = time.time()
db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=os.cpu_count(), dup=dict(min_copies=2, window=1000))
contigs = [db.contig(i) for i in range(db.num_contigs())]
batch = synth.simulate_reads_fast(contigs, a.reads, 150, seed=5, paired=a.paired, inner_mean=300.0, inner_sd=30.0, per_penalty=50.0)
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
parity.feed_reference(g, db)
t1 = time.time()
g.build_index(150)
t2 = time.time()
g.build_duplications(-1, -1, 2, 1000)
print("index build on the device: %.2f s; duplication table: %.2f s" % (t2 - t1, time.time() - t2), flush=True)
if a.host_index:
    t1 = time.time()
    g.build_index(150, threads=os.cpu_count())
    print("index build by the host builder (%d threads): %.2f s" % (os.cpu_count(), time.time() - t1), flush=True)
t1 = time.time()
got = g.align_batch(batch, strict=True)
print("GPU align: %.2f s wall, kernels %.1f ms, aligned %d / %d" % (time.time() - t1, got["stats"][capi.STAT["kernel_ns"]] / 1e6,
      int((np.diff(got["comp_choice_off"])[got["q_comp_off"][:-1]] > 0).sum()), a.reads), flush=True)
t1 = time.time()
want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=os.cpu_count())
print("oracle align (incl. its lazy index build): %.1f s" % (time.time() - t1), flush=True)
parity.assert_same_results(want, got, "large multi-contig")
print("IDENTICAL: %d reads, %d contigs, %d reference bases" % (a.reads, len(ref), sum(len(s) for _, s in ref)))
