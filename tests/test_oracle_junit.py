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
