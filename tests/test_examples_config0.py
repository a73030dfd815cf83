"""BASELINE.json configs[0] on the CPU side: the oracle's SAM / VCF / mutations bodies for examples/ equal the committed fixture
(tests/golden/examples_expected.json, oracle-generated: the reference ships no expected output), and the device core compiled for the host
(emulation harness) returns the same alignments.  The GPU leg is tests/test_gpu_variants.py::test_examples_config0."""
import json
import os
import sys

import parity
import xm_emu
import xm_oracle as xo

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def test_oracle_reproduces_the_examples_fixture():
    import make_examples_expected
    assert make_examples_expected.build() == json.load(open(os.path.join(HERE, "golden", "examples_expected.json")))


def test_examples_through_the_device_core_on_the_host():
    ex = V["examples"]
    db = xo.Oracle([(n, t) for n, t in ex["reference"]], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    batch = parity.batch_from_texts([[t] for _, t in ex["queries"]])
    emu = xm_emu.Emu(ex["params"])
    parity.feed_reference(emu, db)
    emu.build_index(max(len(t) for _, t in ex["queries"]) + 2, threads=1)
    emu.build_duplications(-1, -1, 2, 1000)
    a = db.align_batch(ex["params"], batch)
    b = emu.align_batch(batch, threads=1)
    emu.close()
    parity.assert_same_results(a, b, "examples")
    # query1..5 align, query6-too-different does not (names of examples/queries.fasta)
    per_query = [int(a["comp_choice_off"][a["q_comp_off"][q] + 1] - a["comp_choice_off"][a["q_comp_off"][q]]) for q in range(6)]
    assert per_query[:4] == [1, 1, 1, 1] and per_query[4] >= 1 and per_query[5] == 0
