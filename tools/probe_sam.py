#!/usr/bin/env python3
"""Times xm_format_sam on the bench workload (GPU box): python tools/probe_sam.py [--reads N]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapper_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
a = ap.parse_args()
ref = synth.random_reference(5000000, seed=1)
batch = synth.simulate_reads_fast(ref, a.reads, 150, seed=2)
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
g.set_reference([synth.pack_contig(s) for _, s in ref], [len(s) for _, s in ref])
g.build_index(150)
g.build_duplications(-1, -1, 2, 1000)
names = ["read%07d" % i for i in range(a.reads)]
enc = [n.encode() for n in names]
off = np.zeros(len(enc) + 1, dtype=np.int64); off[1:] = np.cumsum([len(e) for e in enc])
sb = b"".join(enc) + b"\0"
cb = b"contig0\0"; co = np.array([0, 7], dtype=np.int64)
for it in range(3):
    r = C.c_void_p()
    rc = g.L.xm_align_batch(g.h, a.reads, capi._ptr(batch["packed"]), capi._ptr(batch["seq_word_off"]), capi._ptr(batch["seq_len"]), capi._ptr(batch["n_seqs"]),
                            capi._ptr(batch["expected_inner"]), capi._ptr(batch["per_penalty"]), C.byref(r))
    text, n = C.c_char_p(), C.c_int64()
    t0 = time.time()
    g._ok(g.L.xm_format_sam(g.h, r, sb, capi._ptr(off), cb, capi._ptr(co), C.byref(text), C.byref(n)))
    dt = time.time() - t0
    print("xm_format_sam: %d reads, %.1f MB of SAM text in %.1f ms (names H2D + 2 kernels + scan + text D2H) = %.2f M reads/s, %.2f GB/s of text" % (a.reads, n.value / 1e6, dt * 1e3, a.reads / dt / 1e6, n.value / dt / 1e9))
    if it == 2:
        print(C.string_at(text, 400).decode())
    g.L.xm_release_results(r)
