"""Pins tests/variants_oracle.py (the checker of the device count / variant tables) against the reference's own
known-answer tests: T/MutationsWriter_Test.java (7 exact tables, through the oracle aligner), T/MatchDatabase_Test.java (2) and
QVT/VcfWriter_Test.java bodies (hand-built alignments)."""
import json
import os

import numpy as np
import pytest

import parity
import variants_oracle as vo
import xm_oracle as xo
from mapper_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def codes(text):
    return np.array([vo.LETTERS.index(c) for c in text], dtype=np.uint8)


def hand_results(queries):
    """queries: per query a list of components, each a list of choices, each a list of (contig, reversed, blocks[[aS,bS,aL,bL]...])."""
    r = dict(q_comp_off=[0], comp_choice_off=[0], choice_sa_off=[0], sa_block_off=[0], choice_f64=[], sa_f64=[], choice_inner=[], sa_contig=[],
             blocks=[], q_status=[], sa_reversed=[])
    for comps in queries:
        r["q_status"].append(0)
        for choices in comps:
            for sas in choices:
                for contig, rev, blocks in sas:
                    r["sa_contig"].append(contig); r["sa_reversed"].append(rev); r["sa_f64"] += [0.0, 0.0]
                    for b in blocks:
                        r["blocks"] += b
                    r["sa_block_off"].append(r["sa_block_off"][-1] + len(blocks))
                r["choice_sa_off"].append(r["choice_sa_off"][-1] + len(sas)); r["choice_f64"] += [0.0] * 4; r["choice_inner"].append(0)
            r["comp_choice_off"].append(r["comp_choice_off"][-1] + len(choices))
        r["q_comp_off"].append(r["q_comp_off"][-1] + len(comps))
    dt = dict(q_comp_off=np.int64, comp_choice_off=np.int64, choice_sa_off=np.int64, sa_block_off=np.int64, choice_f64=np.float64, sa_f64=np.float64,
              choice_inner=np.int32, sa_contig=np.int32, blocks=np.int32, q_status=np.int32, sa_reversed=np.uint8)
    return {k: np.array(v, dtype=dt[k]) for k, v in r.items()}


@pytest.mark.parametrize("case", V["mutations_cases"], ids=[c["name"] for c in V["mutations_cases"]])
def test_mutations_writer_tables(case):
    db = xo.Oracle([("ref", case["reference"])], sort_by_length=False, dup=case["dup"])
    batch = parity.batch_from_texts([[case["query"]]])
    res = db.align_batch(case["params"], batch)
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    store = vo.Store(contigs, case["end_fraction"])
    vo.accumulate(store, res, synth.unpack_reads(batch), ["query"])
    assert vo.mutations_body(store, vo.Filter(**case["filter"])) == case["expected"]


@pytest.mark.parametrize("case", V["match_database_cases"], ids=[c["name"] for c in V["match_database_cases"]])
def test_match_database_counts(case):
    contigs = [("ref", codes(case["reference"]))]
    sas = [(0, 0, [[0, a["b_start"], a["length"], a["length"]]]) for a in case["alignments"]]
    res = hand_results([[[sas]]])
    store = vo.Store(contigs, case["end_fraction"])
    vo.accumulate(store, res, [[codes(s) for s in case["seqs"]]], ["q%d" % (i + 1) for i in range(len(case["seqs"]))])
    for i in range(len(case["reference"])):
        assert store.position(0, i).count() == 1, i


@pytest.mark.parametrize("case", V["vcf_cases"], ids=[c["name"] for c in V["vcf_cases"]])
def test_vcf_bodies(case):
    contigs = [("contig1", codes(case["reference"]))]
    n = len(case["seq"])
    choices = [[(0, 0, [[0, s, n, n]])] for s in case.get("starts", [0])]
    res = hand_results([[choices]])
    store = vo.Store(contigs, 0.0)
    vo.accumulate(store, res, [[codes(case["seq"])]], ["name1"])
    assert vo.vcf_body(store, include_non_mutations=True, show_support=case["support"]) == case["expected"]


def test_example_choice_is_a_total_order():
    """DirectionalAlignments.betterExample :63-96: longer read, then closer to the read's middle, then earlier position, then the
    lexicographically later name, then the smaller id - the winner must not depend on the order the candidates arrive in."""
    import itertools
    cands = [(vo.Seq("b", 5, codes("ACGTACGTAC")), 4), (vo.Seq("a", 6, codes("ACGTACGTAC")), 4), (vo.Seq("b", 4, codes("ACGTACGTAC")), 4),
             (vo.Seq("c", 1, codes("ACGTACGTAC")), 7), (vo.Seq("c", 2, codes("ACGTACGT")), 4), (vo.Seq("a", 9, codes("ACGTACGTAC")), 6)]
    winners = set()
    for perm in itertools.permutations(cands):
        v = vo.Variant("A")
        for seq, pos in perm:
            if vo.better_example(v, seq, pos):
                v.ex, v.ex_index = seq, pos
        winners.add((v.ex.name, v.ex.id, v.ex_index))
    assert winners == {("b", 4, 4)}
