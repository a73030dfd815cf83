"""Two-GPU test of the in-library NCCL reduction of the count planes and of the sparse variant table (xm_comm_init / xm_counts_reduce:
ncclAllReduce of the planes, all-gather + device sort/reduce-by-key of the variant entries).  Needs >= 2 GPUs; skipped otherwise."""
import threading

import numpy as np
import pytest

import parity
import xm_oracle as xo
from mapper_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_counts_reduce_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = synth.random_reference(120000, seed=201, n_contigs=2)
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batches = [synth.simulate_reads(contigs, 3000, 120, seed=210 + r, sub_rate=0.01, indel_rate=0.002) for r in range(2)]

    def make(dev):
        g = capi.XMapper(synth.DEFAULT_PARAMS, device=dev)
        parity.feed_reference(g, db)
        g.build_index(120)
        g.build_duplications(-1, -1, 2, 1000)
        g.counts_enable(0.1)
        return g

    # reference: one GPU sees both batches
    g = make(0)
    firsts = [0, int(batches[0]["n_seqs"].astype(np.int64).sum())]
    for b, f in zip(batches, firsts):
        g.counts_batch_info(f)
        g.align_batch(b, strict=True)
    want = [g.counts_fetch(c).copy() for c in range(db.num_contigs())]
    want_var = g.variants_fetch()
    assert len(want_var["key"]) > 1000
    g.close()
    # two handles on two GPUs, one batch each, reduced inside the library
    hs = [make(0), make(1)]
    uid = hs[0].comm_unique_id()
    errs = []

    def run(rank):
        try:
            hs[rank].comm_init(2, rank, uid)
            hs[rank].counts_batch_info(firsts[rank])
            hs[rank].align_batch(batches[rank], strict=True)
            hs[rank].counts_reduce()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    for rank in range(2):
        for c in range(db.num_contigs()):
            assert np.array_equal(hs[rank].counts_fetch(c), want[c]), (rank, c)
        have_var = hs[rank].variants_fetch()
        for k in ("key", "count", "ex_gid", "ex_rev", "ex_index"):
            assert np.array_equal(have_var[k], want_var[k]), (rank, k)
    [h.close() for h in hs]
