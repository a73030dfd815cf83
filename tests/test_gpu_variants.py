"""Row f1 of SURVEY.md §8 on the GPU: the device's count planes + sparse variant table (xm_counts_* / xm_variants_fetch) against the
Python restatement of QuickVariants' MatchDatabase (tests/variants_oracle.py, pinned by the reference's MutationsWriter / MatchDatabase /
VcfWriter KATs), entry by entry and through the VCF and mutations BODIES; plus BASELINE.json configs[0] (examples/) end to end."""
import json
import os

import numpy as np
import pytest

import parity
import sam_oracle
import variants_oracle as vo
import xm_oracle as xo
from mapper_b200 import capi, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def name_ranks(names):
    """A key that orders sequences like their names (what xm_counts_batch_info asks the host for)."""
    order = {n: i for i, n in enumerate(sorted(set(names)))}
    return np.array([order[n] for n in names], dtype=np.int64)


def device_store(g, contigs, end_fraction, all_reads, all_names):
    """Rebuilds the oracle's Store from what the device accumulated.  all_reads / all_names: every sequence of the run by global id."""
    planes = [g.counts_fetch(c) for c in range(len(contigs))]
    t = g.variants_fetch()
    off = np.concatenate([[0], np.cumsum([len(c) for _, c in contigs])])
    contig = np.searchsorted(off, t["gpos"], side="right") - 1
    t["contig"] = contig
    t["pos"] = t["gpos"] - off[contig]

    def lookup(gid, rev):
        codes = all_reads[gid]
        return vo.Seq(all_names[gid] + ("-rev" if rev else ""), gid, vo.COMP[codes[::-1]] if rev else codes)
    return vo.Store.from_device(contigs, end_fraction, planes, t, lookup), t


def flat(reads):
    return [m for q in reads for m in q]


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_variant_table_and_bodies_vs_oracle(paired):
    from test_emu_parity import ambiguate
    ref = synth.random_reference(24000, seed=301, n_contigs=3, repeat_fraction=0.15, repeat_len=(150, 600), repeat_divergence=0.0)
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
    parity.feed_reference(g, db)
    g.build_index(100)
    g.build_duplications(-1, -1, 2, 1000)
    g.counts_enable(0.1)
    store = vo.Store(contigs, 0.1)
    all_reads, all_names = [], []
    rng = np.random.default_rng(7)
    universe = sorted("r%d" % i for i in range(40))
    for it in range(3):
        batch = ambiguate(synth.simulate_reads(contigs, 2500, 100, seed=302 + 3 * it + paired, sub_rate=0.02, indel_rate=0.006, paired=paired,
                                               inner_mean=60.0, inner_sd=40.0), 303 + it, 0.004)
        reads = synth.unpack_reads(batch)
        n_seq = sum(len(q) for q in reads)
        # few distinct names, so that the example choice regularly reaches the name and id tie-breaks
        names = ["r%d" % rng.integers(0, 40) for _ in range(n_seq)]
        first = len(all_names)
        all_reads += flat(reads); all_names += names
        g.counts_batch_info(first, np.array([universe.index(n) for n in names], dtype=np.int64))  # ranks comparable across batches
        got = g.align_batch(batch, strict=True)
        vo.accumulate(store, got, reads, names, first_seq_id=first)
    dstore, t = device_store(g, contigs, 0.1, all_reads, all_names)
    want = store.variant_table()
    have = dstore.variant_table()
    assert len(want) > 3000 and any(w[4] >= 0 for w in want) and any(w[5] == 5 for w in want)
    assert have == want
    assert np.all(np.diff(t["key"].astype(np.uint64)) > 0)  # sorted, unique keys
    for c in range(len(contigs)):
        for region in range(2):
            for d in range(2):
                assert np.array_equal(dstore.d[c][region][d].ref_counts, store.d[c][region][d].ref_counts)
    for flt in (vo.Filter(), vo.Filter.default()):
        assert vo.mutations_body(dstore, flt) == vo.mutations_body(store, flt)
    body = vo.vcf_body(store)
    assert vo.vcf_body(dstore) == body and body.count("\n") > 20000
    g.close()


def test_examples_config0():
    """BASELINE.json configs[0]: examples/reference.fasta + examples/queries.fasta with the defaults of M/Mapper.java:409-453 as
    examples/test.sh:14 runs them (--out-sam --out-vcf; single-end).  Alignments, SAM text, count planes, variant table, VCF and
    mutations bodies: device vs oracle, and vs the committed oracle-generated fixture tests/golden/examples_expected.json."""
    ex = V["examples"]
    db = xo.Oracle([(n, t) for n, t in ex["reference"]], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    assert [n for n, _ in contigs] == ["contig3", "contig1", "contig2"]  # Mapper.sortAndComplementReference :1151-1172
    names = [n for n, _ in ex["queries"]]
    batch = parity.batch_from_texts([[t] for _, t in ex["queries"]])
    g = capi.XMapper(ex["params"], device=0)
    parity.feed_reference(g, db)
    g.build_index(max(len(t) for _, t in ex["queries"]) + 2)
    g.build_duplications(-1, -1, 2, 1000)
    g.counts_enable(0.1)
    g.counts_batch_info(0, name_ranks(names))
    got, sam = g.align_batch_sam(batch, names, [n for n, _ in contigs], strict=True)
    want = db.align_batch(ex["params"], batch)
    parity.assert_same_results(want, got, "examples")
    reads = synth.unpack_reads(batch)
    store = vo.Store(contigs, 0.1)
    vo.accumulate(store, want, reads, names)
    dstore, _ = device_store(g, contigs, 0.1, flat(reads), names)
    assert dstore.variant_table() == store.variant_table()
    vcf, mut = vo.vcf_body(dstore), vo.mutations_body(dstore, vo.Filter.default())
    assert vcf == vo.vcf_body(store) and mut == vo.mutations_body(store, vo.Filter.default())
    golden = json.load(open(os.path.join(HERE, "golden", "examples_expected.json")))
    assert sam == golden["sam_body"] and sam == sam_oracle.format_sam(want, batch, names, [n for n, _ in contigs])
    assert vcf == golden["vcf_body"] and mut == golden["mutations_body"]
    aligned = [int(got["comp_choice_off"][got["q_comp_off"][q] + 1] - got["comp_choice_off"][got["q_comp_off"][q]]) for q in range(len(names))]
    assert aligned == golden["choices_per_query"]
    g.close()
