#!/usr/bin/env python3
"""Runs mixed batches through the probe build of the library (libxmapper_b200_dbg.so, -DXM_DBG_UNIFORM) and checks that no query
reports a split warp or a lane-dependent "uniform" value (statuses -7777, -8000..-8002, <= -100000, -2000-line, -3000-line), and that
the results still equal the oracle's.  The per-query code runs on warp-shared state with all 32 lanes: it is only correct while the
warp is converged.  usage (GPU box): XM_LIB_PATH=mapper_b200/libxmapper_b200_dbg.so python tools/check_convergence.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import xm_oracle as xo  # noqa: E402
from mapper_b200 import capi, synth  # noqa: E402
from test_emu_parity import ambiguate  # noqa: E402

assert "dbg" in os.path.basename(capi.LIB_PATH), "set XM_LIB_PATH to the probe build (libxmapper_b200_dbg.so)"
ref = synth.random_reference(300000, seed=131, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 1000))
db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
contigs = [db.contig(i) for i in range(db.num_contigs())]
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
parity.feed_reference(g, db)
g.build_index(1000)
g.build_duplications(-1, -1, 2, 1000)
batches = [
    ("ambiguous reads (escalate in place)", ambiguate(synth.simulate_reads(contigs, 8000, 150, seed=132, sub_rate=0.01, indel_rate=0.002), 133, 0.01)),
    ("paired reads", synth.simulate_reads(contigs, 6000, 150, seed=134, sub_rate=0.02, indel_rate=0.003, paired=True)),
    ("1 kbp reads", synth.simulate_reads(contigs, 600, 1000, seed=135, sub_rate=0.01, indel_rate=0.005)),
]
bad = 0


def run(g, db, params, name, batch):
    global bad
    for rep in range(2):
        got = g.align_batch(batch)
        vals, cnt = np.unique(got["q_status"], return_counts=True)
        if len(vals) != 1 or vals[0] != 0:
            bad += 1
            print("SPLIT", name, dict(zip(vals.tolist(), cnt.tolist())))
    want = db.align_batch(params, batch, threads=8)
    parity.assert_same_results(want, got, name)
    print("ok:", name)


for name, batch in batches:
    run(g, db, synth.DEFAULT_PARAMS, name, batch)
g.close()
# reads with long ambiguous tails, short reads, and two non-default penalty models (other paths through the cascade)
many_n = synth.simulate_reads(contigs, 1500, 150, seed=136, sub_rate=0.01, indel_rate=0.002)
codes = [q[0].copy() for q in synth.unpack_reads(many_n)]
for i, r in enumerate(codes):
    if i % 3 == 0:
        r[-(20 + i % 60):] = 15
many_n = synth.batch_from_reads(codes)
short = synth.simulate_reads(contigs, 3000, 40, seed=137, sub_rate=0.02, indel_rate=0.005)
from test_emu_parity import variant_params  # noqa: E402
for vname in ("cheap-indels", "loose-error-rate"):
    p = variant_params(vname)
    g = capi.XMapper(p, device=0)
    parity.feed_reference(g, db)
    g.build_index(150)
    g.build_duplications(-1, -1, 2, 1000)
    run(g, db, p, vname + ": reads with N tails", many_n)
    run(g, db, p, vname + ": 40 bp reads", short)
    g.close()
sys.exit(1 if bad else 0)
