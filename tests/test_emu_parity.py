"""Device-core logic (compiled for the host by the emulation harness) vs the oracle, bit for bit.
These run without a GPU; the same comparisons run against the real CUDA library in test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import parity
import xm_emu
import xm_oracle as xo
from mapper_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def run_both(oracle_db, params, batch, window, max_used, upload_index=True, upload_dups=True, dup=None):
    emu = xm_emu.Emu(params)
    parity.feed_from_oracle(emu, oracle_db, max_used, window, upload_index, upload_dups)
    if not upload_index:
        emu.build_index(max_used, threads=2)
    if not upload_dups:
        d = dup or {}
        emu.build_duplications(d.get("min_len", -1), d.get("max_len", -1), d.get("min_copies", 2), window)
    a = oracle_db.align_batch(params, batch, threads=1)
    b = emu.align_batch(batch, threads=1)
    emu.close()
    return a, b


@pytest.mark.parametrize("case", V["api_cases"], ids=[c["name"] for c in V["api_cases"]])
def test_junit_api_cases(case):
    db = xo.Oracle([("reference-0", case["reference"])], dup=dict(min_copies=2, window=1))
    batch = parity.batch_from_texts([case["seqs"]], [case["expected_inner"]], [case["per_penalty"]])
    max_used = max(len(s) for s in case["seqs"]) + 2
    a, b = run_both(db, case["params"], batch, 1, max_used)
    parity.assert_same_results(a, b, case["name"])
    # and the reference's own expectation
    nchoice = b["comp_choice_off"][1] - b["comp_choice_off"][0] if len(b["comp_choice_off"]) == 2 else 0
    assert nchoice == case["expect"]["count"]


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_random_reads_small_reference(paired):
    ref = synth.random_reference(300000, seed=11, n_contigs=3, repeat_fraction=0.08, repeat_len=(200, 1500))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = synth.simulate_reads(contigs, 3000, 150, seed=12 + paired, sub_rate=0.02, indel_rate=0.003, paired=paired)
    a, b = run_both(db, synth.DEFAULT_PARAMS, batch, 1000, 150)
    assert (a["q_status"] == 0).all()
    parity.assert_same_results(a, b, "random %s" % ("paired" if paired else "single"))


def test_library_index_builder_matches_host_tables():
    ref = synth.random_reference(200000, seed=21, n_contigs=2, repeat_fraction=0.1, repeat_len=(100, 800))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    built = db.build_through(120)
    emu = xm_emu.Emu(synth.DEFAULT_PARAMS)
    parity.feed_reference(emu, db)
    emu.build_index(120, threads=3)
    mi, mb = emu.index_info()
    assert mi == db.min_interesting()
    for n in range(1, min(built, mb) + 1):
        t0, t1 = db.table(n), emu.get_index_length(n)
        assert t0["capacity"] == t1["capacity"] and t0["max_count"] == t1["max_count"], n
        assert np.array_equal(t0["overfull"], t1["overfull"]), n
        assert np.array_equal(t0["positions"], t1["positions"]), n
        keep = t0["overfull"] == 0
        assert np.array_equal(np.diff(t0["offsets"])[keep], np.diff(t1["offsets"])[keep]), n
    # duplication keys
    db.detect_duplications()
    emu.build_duplications(-1, -1, 2, 1000)
    for c in range(db.num_contigs()):
        assert np.array_equal(db.dup_starts(c), emu.get_duplications(c)), c
    # the merge the device scan feeds (forward-strand blocks, units merged by different threads) gives the same table
    for window, lo, hi, copies in ((1000, -1, -1, 2), (1, -1, -1, 2), (100, 12, 30, 3)):
        emu.build_duplications(lo, hi, copies, window)
        want = [emu.get_duplications(c).copy() for c in range(db.num_contigs())]
        emu.build_duplications(lo, hi, copies, window, via_merge=True)
        for c in range(db.num_contigs()):
            assert np.array_equal(want[c], emu.get_duplications(c)), (window, c)
        assert sum(len(w) for w in want) > 0
    emu.close()


def test_positions_past_2_32():
    """Global positions wider than 32 bits (QV/SequenceDatabase.java:69-74 sizes them by the forward + reverse size; KAT
    T/PackedMap_Test.testLargeReferenceSize): the reference is placed 2^33 + 12345 bases into the position space, as if a reference of
    that size preceded it.  The host builder's tables are the unbiased ones shifted by the bias, the duplication table is unchanged,
    uploaded 64-bit tables work, and the alignments are bit-identical to the oracle's."""
    bias = (1 << 33) + 12345
    ref = synth.random_reference(120000, seed=211, n_contigs=3, repeat_fraction=0.1, repeat_len=(100, 600))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = synth.simulate_reads(contigs, 600, 120, seed=212, sub_rate=0.02, indel_rate=0.003, paired=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=2)
    plain = xm_emu.Emu(synth.DEFAULT_PARAMS)
    parity.feed_reference(plain, db)
    plain.build_index(120, threads=2)
    plain.build_duplications(-1, -1, 2, 1000)
    wide = xm_emu.Emu(synth.DEFAULT_PARAMS)
    wide.set_position_bias(bias)
    parity.feed_reference(wide, db)
    wide.build_index(120, threads=2)
    wide.build_duplications(-1, -1, 2, 1000)
    mi, mb = wide.index_info()
    assert (mi, mb) == plain.index_info()
    tables = []
    for n in range(1, mb + 1):
        t0, t1 = plain.get_index_length(n), wide.get_index_length(n, wide=True)
        assert t0["capacity"] == t1["capacity"] and np.array_equal(t0["offsets"], t1["offsets"]) and np.array_equal(t0["overfull"], t1["overfull"]), n
        assert np.array_equal(t0["positions"].astype(np.uint64) + np.uint64(bias), t1["positions"]), n
        tables.append(t1)
    for c in range(db.num_contigs()):
        assert np.array_equal(plain.get_duplications(c), wide.get_duplications(c)), c
    parity.assert_same_results(want, wide.align_batch(batch, threads=1), "biased positions, host-built index")
    db.detect_duplications()
    up = xm_emu.Emu(synth.DEFAULT_PARAMS)   # the Java host's path: tables uploaded with 64-bit positions
    up.set_position_bias(bias)
    parity.feed_reference(up, db)
    for t in tables:
        up.set_index_length(t, wide=True)
    up.finish_index(mi, mb)
    for c in range(db.num_contigs()):
        up.set_duplications(1000, db.dup_granularity(), c, plain.get_duplications(c))
    parity.assert_same_results(want, up.align_batch(batch, threads=1), "biased positions, uploaded index")
    for e in (plain, wide, up):
        e.close()


def ambiguate(batch, seed, rate, codes=(15, 15, 15, 5, 10, 3, 12, 7)):
    """Replaces a fraction of the query bases by IUPAC-ambiguous codes that still contain the original base (N mostly, some
    two- and three-base codes), in the QV 4-bit packing the batch carries."""
    rng = np.random.default_rng(seed)
    packed = batch["packed"].copy()
    off = batch["seq_word_off"]
    for s, ln in enumerate(batch["seq_len"]):
        if ln < 1:
            continue
        k = rng.binomial(int(ln), rate)
        if s % 7 == 0:
            k += 2  # some reads with several, including adjacent ones
        for pos in rng.integers(0, int(ln), size=k).tolist() + ([1, 2] if s % 7 == 0 and ln > 3 else []):
            wi = int(off[s]) + (pos >> 2)
            sh = (pos & 3) << 2
            old = (int(packed[wi]) >> sh) & 15
            code = int(codes[int(rng.integers(0, len(codes)))]) | old
            packed[wi] = np.uint16((int(packed[wi]) & ~(15 << sh)) | (code << sh))
    out = dict(batch)
    out["packed"] = packed
    return out


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_reads_with_ambiguous_bases(paired):
    """IUPAC-ambiguous QUERY bases: MultiHashBlocks in the query pyramid (M/HashBlock_ParentRow.java:69-191), stepped past by the
    seed walk (M/HashBlockPath.java:130-140), scored with the ambiguity penalty (M/AlignmentParameters.java:156-180)."""
    ref = synth.random_reference(200000, seed=31, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 1000))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = ambiguate(synth.simulate_reads(contigs, 1500, 120, seed=32 + paired, sub_rate=0.01, indel_rate=0.002, paired=paired), 33, 0.01)
    a, b = run_both(db, synth.DEFAULT_PARAMS, batch, 1000, 120)
    assert (b["q_status"] == 0).all()
    parity.assert_same_results(a, b, "ambiguous reads %s" % ("paired" if paired else "single"))


PARAM_VARIANTS = {
    "no-gapmers": dict(enable_gapmers=0),
    "cheap-indels": dict(ins_start=0.8, ins_ext=0.3, del_start=0.7, del_ext=0.25),
    "loose-error-rate": dict(max_error_rate=0.2),
    "no-span-one-match": dict(max_penalty_span=0.0, max_num_matches=1),
    "costly-mutation": dict(mutation=2.0, ambiguity=0.3, unaligned=0.25, max_penalty_span=1.5),
}


def variant_params(name):
    p = dict(synth.DEFAULT_PARAMS)
    p.update(PARAM_VARIANTS[name])
    return p


@pytest.mark.parametrize("name", sorted(PARAM_VARIANTS), ids=sorted(PARAM_VARIANTS))
def test_parameter_variants(name):
    """Non-default AlignmentParameters (M/AlignmentParameters.java:6-37, --no-gapmers M/Mapper.java:51): the penalty model is data, not code."""
    p = variant_params(name)
    ref = synth.random_reference(150000, seed=41, n_contigs=2, repeat_fraction=0.08, repeat_len=(150, 800))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, gapmers=bool(p.get("enable_gapmers", 1)), dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    for paired in (False, True):
        batch = synth.simulate_reads(contigs, 700, 100, seed=43 + paired, sub_rate=0.02, indel_rate=0.004, paired=paired, inner_mean=150.0, inner_sd=20.0)
        a, b = run_both(db, p, batch, 1000, 100)
        parity.assert_same_results(a, b, "%s paired=%s" % (name, paired))
