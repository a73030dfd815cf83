"""Pins the oracle (CPU restatement) against the reference's own JUnit known-answer tests (tests/golden/junit_vectors.json)."""
import json
import os

import pytest

import xm_oracle as xo

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def api_db(reference, dup=None):
    # Api.newDatabase(referenceText): one contig "reference-0" (+ RC), DuplicationDetector(min, max, 2 copies, window 1) (M/Api.java:41-69)
    return xo.Oracle([("reference-0", reference)], sort_by_length=False, dup=dup or dict(min_copies=2, window=1))


@pytest.mark.parametrize("case", V["api_cases"], ids=[c["name"] for c in V["api_cases"]])
def test_api_alignOnce(case):
    db = api_db(case["reference"])
    r = db.align(case["params"], case["seqs"], case["expected_inner"], case["per_penalty"])
    e = case["expect"]
    top = r["components"][0] if len(r["components"]) == 1 else []  # QueryAlignments.getTopLevelAlignments
    assert len(top) == e["count"], json.dumps(r, indent=1)
    if "aligned_b0" in e:
        assert top[0]["seqs"][0]["aligned_b"] == e["aligned_b0"]
    if "start_b" in e:
        got = [[s["start_b"] for s in ch["seqs"]] for ch in top]
        assert got == e["start_b"]


@pytest.mark.parametrize("case", V["path_aligner_cases"], ids=[c["name"] for c in V["path_aligner_cases"]])
def test_path_aligner(case):
    r = xo.path_aligner(case["params"], case["a"], case["b"], case["penalty"], case["penalty"])
    assert r is not None
    assert r["penalty"] == case["penalty"]  # exact double, as the reference test (T/PathAligner_Test.java:68)
    assert r["aligned_a"] == case["aligned_a"]
    assert r["aligned_b"] == case["aligned_b"]


@pytest.mark.parametrize("case", V["hashblock_aligner_cases"], ids=[c["name"] for c in V["hashblock_aligner_cases"]])
def test_hashblock_aligner(case):
    r = xo.hashblock_aligner(case["params"], case["a"], case["b"], case["penalty"], case["penalty"])
    assert r is not None
    assert r["aligned_a"] == case["aligned_a"]
    assert r["aligned_b"] == case["aligned_b"]
    assert abs(r["penalty"] - case["penalty"]) <= 0.000001  # T/HashBlockAligner_Test.java:76


@pytest.mark.parametrize("case", V["counting_path_cases"], ids=[c["name"] for c in V["counting_path_cases"]])
def test_counting_path(case):
    # new SequenceDatabase(reference, true); new HashBlock_Database(sequenceDatabase) (T/Counting_HashBlockPath_Test.java:64-75)
    db = xo.Oracle([("reference", case["reference"])])
    offs = db.counting_path(case["params"], case["query"], len(case["query"]))
    e = case["expect"]
    if "num_offsets" in e:
        assert len(offs) == e["num_offsets"], offs
    if "contains_offset" in e:
        assert e["contains_offset"] in [o[1] for o in offs], offs


@pytest.mark.parametrize("case", V["paths_counter_cases"], ids=[c["name"] for c in V["paths_counter_cases"]])
def test_paths_counter(case):
    db = xo.Oracle([("ref", case["reference"])])
    m = db.paths_counter(case["params"], case["seq1"], case["seq2"], 10, 20)
    e = case["expect"]
    assert len(m) == e["count"], m
    if "inner" in e:
        assert m[0]["inner"] == e["inner"] and m[0]["across"] == e["across"], m


@pytest.mark.parametrize("text", V["symmetry_cases"])
def test_hash_symmetry(text):
    assert xo.hash_symmetry(text) > 0


# ---- T/SamWriter_Test.java: exact SAM bodies (pins tests/sam_oracle.py, the checker of xm_format_sam) ----
@pytest.mark.parametrize("case", V["sam_cases"], ids=[c["name"] for c in V["sam_cases"]])
def test_sam_bodies(case):
    import parity
    import sam_oracle
    db = xo.Oracle([(case["ref_name"], case["reference"])], sort_by_length=False, dup=case["dup"])
    batch = parity.batch_from_texts([case["seqs"]], [case["expected_inner"]], [case["per_penalty"]])
    r = db.align_batch(case["params"], batch)
    assert sam_oracle.format_sam(r, batch, case["names"], [case["ref_name"]]) == case["expected_sam"]


def test_java_float_formatting():
    import sam_oracle
    assert sam_oracle.format_number(0.0) == "f:0.0"
    assert sam_oracle.format_number(1.5) == "f:-1.5"
    assert sam_oracle.format_number(0.1 / 3) == "f:-0.0333"
    assert sam_oracle.format_number(12.3456789) == "f:-12.3457"
    assert sam_oracle.format_number(0.0001) == "f:-1.0E-4"
    assert sam_oracle.format_number(1500.12341) == "f:-1500.1234"
    assert sam_oracle.java_float_str(1.0e7) == "1.0E7"
    assert sam_oracle.java_float_str(123456.7) == "123456.7"


# ---- T/BasepairsTest.java:9-47 (through the oracle's StraightAligner penalty: one base pair, AmbiguityPenalty 3, MutationPenalty 100) ----
@pytest.mark.parametrize("case", V["basepair_cases"], ids=[c["q"] + c["r"] for c in V["basepair_cases"]])
def test_basepair_penalties(case):
    import ctypes as C
    p = dict(mutation=100.0, ins_start=1.0, ins_ext=1.0, del_start=1.0, del_ext=1.0, max_error_rate=1000.0, unaligned=1.0, ambiguity=3.0, max_penalty_span=0.0)
    want = {"mutation": 100.0, "ambiguity": 3.0, "ambiguity/3": 3.0 / 3}[case["penalty"]]
    L = xo.lib()
    L.xo_test_base_penalty.restype = C.c_double
    cp = xo.make_params(p)
    for q, r in ((case["q"], case["r"]), (case["r"], case["q"])):  # the test checks both argument orders
        assert L.xo_test_base_penalty(C.byref(cp), q.encode(), r.encode()) == want


# ---- T/AncestryDetector_Test.java:10-91: seven exact inferred-ancestor strings ----
@pytest.mark.parametrize("case", V["ancestry_cases"], ids=[c["name"] for c in V["ancestry_cases"]])
def test_ancestry_detector(case):
    from mapper_b200 import synth
    db = xo.Oracle([("ref", case["reference"])], sort_by_length=False, dup=dict(min_copies=3, window=1))
    anc = db.infer_ancestors(case["threshold"], verify=True)
    assert len(anc) == 1 and anc[0][0] == "ref-anc"
    assert synth.codes_to_text(anc[0][1]) == case["expected"]


# ---- T/HashBlockDatabase_Test.java:14-27: the tables do not depend on the order in which the hashing jobs run ----
def test_index_independent_of_job_order():
    import numpy as np
    from mapper_b200 import synth
    contigs = [("contig1", "ACCCCCCC"), ("contig2", "CTTTTTTT")]
    a = xo.Oracle(contigs, min_interesting=1, max_short=1, threads=1)
    b = xo.Oracle(contigs, min_interesting=1, max_short=1, threads=3)
    assert a.build_through(100) == b.build_through(100)
    for n in range(1, a.max_built() + 1):
        ta, tb = a.table(n), b.table(n)
        assert ta["capacity"] == tb["capacity"] and np.array_equal(ta["offsets"], tb["offsets"]) and np.array_equal(ta["positions"], tb["positions"])
    ref = synth.random_reference(120000, seed=3, n_contigs=4, repeat_fraction=0.1, repeat_len=(100, 500))
    texts = [(n, synth.codes_to_text(s)) for n, s in ref]
    a, b = xo.Oracle(texts, sort_by_length=True, threads=1), xo.Oracle(texts, sort_by_length=True, threads=5)
    a.build_through(80); b.build_through(80)
    for n in range(1, 81):
        ta, tb = a.table(n), b.table(n)
        assert np.array_equal(ta["offsets"], tb["offsets"]) and np.array_equal(ta["positions"], tb["positions"]) and np.array_equal(ta["overfull"], tb["overfull"])


# ---- T/MultiHashBlock_Test.java:13-88: ambiguous letters expand into the specific base pairs ----
def _add_ns(text, k):  # addAmbiguities :178-192
    if k < 1:
        return [text]
    if k > len(text):
        return []
    return ["N" + t for t in _add_ns(text[1:], k - 1)] + [text[0] + t for t in _add_ns(text[1:], k)]


@pytest.mark.parametrize("text,max_n", [("A", 1), ("AAA", 3), ("AAAAAAAAAAAAAAA", 3), ("TTATGC", 1)])
def test_multi_hash_block_expanding_ns(text, max_n):
    L = xo.lib()
    if L.xo_test_multi_expand(text.encode(), text.encode()) == -1:
        return  # "We don't have a hashblock that spans the entire sequence" (:93-96)
    n = 0
    for k in range(max_n + 1):
        for amb in _add_ns(text, k):
            assert L.xo_test_multi_expand(text.encode(), amb.encode()) == 1, amb
            n += 1
    assert n >= 2


PARTIAL = [("AAA", "ARA"), ("GGG", "GRG"), ("CCC", "CYC"), ("TTT", "TYT"), ("AAA", "AWA"), ("TTT", "TWT"), ("CCC", "CSC"), ("GGG", "GSG"),
           ("GGG", "GKG"), ("TTT", "TKT"), ("AAA", "AMA"), ("CCC", "CMC"), ("AAA", "ADA"), ("GGG", "GDG"), ("TTT", "TDT"), ("AAA", "AVA"),
           ("CCC", "CVC"), ("GGG", "GVG"), ("AAA", "AHA"), ("CCC", "CHC"), ("TTT", "THT"), ("CCC", "CBC"), ("GGG", "GBG"), ("TTT", "TBT"),
           ("AAAAAA", "ARRRRA")]  # checkPartialAmbiguity :34-78, checkManyPartialAmbiguities :82-85


@pytest.mark.parametrize("text,amb", PARTIAL, ids=[a for _, a in PARTIAL])
def test_multi_hash_block_partial_ambiguity(text, amb):
    assert xo.lib().xo_test_multi_expand(text.encode(), amb.encode()) == 1


# ---- T/PackedMap_Test.java:14-48, T/SequenceDatabase_Test.java:17-41: positions beyond 2^31 ----
def test_global_positions_beyond_2_to_31():
    """The reference's position space is 64-bit (QV/SequenceDatabase.java:69-86,170-209).  The oracle's encode/decode is exercised with 8
    sequences of 2^31 bases (the PackedMap_Test shape) and 16 of 2^30 (SequenceDatabase_Test) without materialising them; the PRODUCT
    stores 32-bit positions and refuses references with 2N >= 2^32 (xm_set_reference) - documented limit, DESIGN.md."""
    import ctypes as C
    L = xo.lib()
    L.xo_test_position_roundtrip.restype = C.c_int64
    for n_seq, length in ((8, 2 ** 31), (16, 2 ** 30), (2 ** 13, 2 ** 21)):
        assert L.xo_test_position_roundtrip(n_seq, C.c_int64(length)) == 4 * n_seq
