import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import xm_oracle as xo, parity
from mapper_b200 import capi, synth
from test_emu_parity import ambiguate
ref = synth.random_reference(300000, seed=131, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 1000))
db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
contigs = [db.contig(i) for i in range(db.num_contigs())]
batch = ambiguate(synth.simulate_reads(contigs, 8000, 150, seed=132, sub_rate=0.01, indel_rate=0.002, paired=False), 133, 0.01)
g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
parity.feed_reference(g, db); g.build_index(150); g.build_duplications(-1, -1, 2, 1000)
try:
    got = g.align_batch(batch)
    st = got["q_status"]
    vals, cnt = np.unique(st, return_counts=True)
    print("q_status values:", dict(zip(vals.tolist(), cnt.tolist())))
except Exception as e:
    print("EXC", e)
