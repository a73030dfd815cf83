"""The BASELINE.json configs at sizes the oracle finishes in well under a minute each, through the C ABI on the GPU, bit for bit against
the oracle: configs[1] (single-end 150 bp), configs[2] (paired 2x150, --spacing 300 50), configs[3] shape (50 contigs, repeat families,
duplication logic), configs[4] as written (10 kbp reads -> --split-queries-past-size 1000 pieces, 1 % substitutions + 0.5 % indels,
>= 3-copy repeat families).  configs[0] is tests/test_gpu_variants.py::test_examples_config0."""
import os

import numpy as np
import pytest

import parity
import xm_oracle as xo
from mapper_b200 import capi, synth

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4


def setup(ref, max_used):
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=THREADS, dup=dict(min_copies=2, window=1000))
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
    parity.feed_reference(g, db)
    g.build_index(max_used)                  # device index builder
    g.build_duplications(-1, -1, 2, 1000)
    return db, g, [db.contig(i) for i in range(db.num_contigs())]


def check(db, g, batch, what):
    got = g.align_batch(batch, strict=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=THREADS)
    parity.assert_same_results(want, got, what)
    return got


def test_config1_single_end_5mbp():
    db, g, contigs = setup(synth.random_reference(5000000, seed=1), 150)
    got = check(db, g, synth.simulate_reads_fast(contigs, 120000, 150, seed=2), "configs[1]")
    aligned = int((np.diff(got["comp_choice_off"])[got["q_comp_off"][:-1]] > 0).sum())
    assert aligned > 0.99 * 120000
    g.close()


def test_config2_paired_end_spacing_300_50():
    db, g, contigs = setup(synth.random_reference(5000000, seed=1), 150)
    batch = synth.simulate_reads_fast(contigs, 40000, 150, seed=3, paired=True, inner_mean=300.0, inner_sd=30.0, per_penalty=50.0)
    got = check(db, g, batch, "configs[2]")
    two_mates = int((np.diff(got["choice_sa_off"]) == 2).sum())
    assert two_mates > 0.95 * 40000     # properly paired: one QueryAlignment with both SequenceAlignments
    g.close()


def test_config3_shape_50_contigs_with_repeats():
    ref = synth.random_reference(30000000, seed=4, n_contigs=50, repeat_fraction=0.05, repeat_copies=(2, 4), repeat_len=(1000, 5000))
    db, g, contigs = setup(ref, 150)
    assert len(contigs) == 50 and sum(len(g.get_duplications(c)) for c in range(50)) > 100   # the duplication logic has something to do
    got = check(db, g, synth.simulate_reads_fast(contigs, 100000, 150, seed=5), "configs[3] shape")
    multi = int((np.diff(got["comp_choice_off"]) > 1).sum())
    assert multi > 50                   # reads inside repeat copies report several choices
    g.close()


def test_config4_long_reads_split_past_1000():
    ref = synth.random_reference(3000000, seed=6, n_contigs=10, repeat_fraction=0.06, repeat_copies=(3, 6), repeat_len=(1000, 5000))
    db, g, contigs = setup(ref, 1000)
    long_batch = synth.simulate_reads(contigs, 60, 10000, seed=7, sub_rate=0.01, indel_rate=0.005)
    long_reads = [q[0] for q in synth.unpack_reads(long_batch)]
    pieces, parent = synth.split_queries(long_reads, 1000)
    # M/SequenceSplitter.java:16,39-41: (len - 1) // 1000 + 1 pieces per read, none longer than 1000, lengths differing by at most one
    assert len(pieces) == sum((len(r) - 1) // 1000 + 1 for r in long_reads) >= 590 and max(len(p) for p in pieces) <= 1000 and min(len(p) for p in pieces) >= 900
    assert sum(len(p) for p in pieces) == sum(len(r) for r in long_reads) and parent[-1] == 59
    got = check(db, g, synth.batch_from_reads(pieces), "configs[4]")
    aligned = int((np.diff(got["comp_choice_off"])[got["q_comp_off"][:-1]] > 0).sum())
    assert aligned > 500
    # a ragged split: 2500 bases past size 1000 -> 3 pieces of 833 / 833 / 834
    rag, _ = synth.split_queries([long_reads[0][:2500]], 1000)
    assert [len(p) for p in rag] == [833, 833, 834]
    check(db, g, synth.batch_from_reads(rag), "configs[4] ragged pieces")
    g.close()


def test_config4_with_inferred_ancestors():
    """configs[4] with --infer-ancestors: the aligner works on the inferred-ancestor reference (IUPAC unions in the repeat copies,
    M/Mapper.java:675-681); index (MultiHashBlock fan-out) and duplication table built by the library; 10 kbp reads split past 1000."""
    ref = synth.random_reference(2000000, seed=6, n_contigs=10, repeat_fraction=0.08, repeat_copies=(3, 6), repeat_len=(1000, 5000))
    db, changed = parity.inferred_ancestor_oracle(ref, synth.DEFAULT_PARAMS, threads=THREADS)
    assert changed > 500
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
    parity.feed_reference(g, db)
    g.build_index(1000, threads=THREADS)
    g.build_duplications(-1, -1, 2, 1000)
    long_batch = synth.simulate_reads(ref, 40, 10000, seed=8, sub_rate=0.01, indel_rate=0.005)
    pieces, _ = synth.split_queries([q[0] for q in synth.unpack_reads(long_batch)], 1000)
    check(db, g, synth.batch_from_reads(pieces), "configs[4] --infer-ancestors")
    g.close()
