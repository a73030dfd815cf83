"""Generates tests/golden/examples_expected.json: what the ORACLE (oracle/ + tests/sam_oracle.py + tests/variants_oracle.py) produces for
BASELINE.json configs[0] - examples/reference.fasta + examples/queries.fasta (transcribed in junit_vectors.json:examples) with the defaults of
M/Mapper.java:409-453, as examples/test.sh:14 runs them.  The reference ships no expected output for the examples (examples/.gitignore) and no
JVM is available, so this fixture is oracle-generated, not reference-generated: it freezes the oracle's answer so that a later change to the
oracle or the device shows up as a diff.  Run from the repo root: python tests/golden/make_examples_expected.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import parity  # noqa: E402
import sam_oracle  # noqa: E402
import variants_oracle as vo  # noqa: E402
import xm_oracle as xo  # noqa: E402
from mapper_b200 import synth  # noqa: E402


def build():
    ex = json.load(open(os.path.join(HERE, "junit_vectors.json")))["examples"]
    db = xo.Oracle([(n, t) for n, t in ex["reference"]], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    names = [n for n, _ in ex["queries"]]
    batch = parity.batch_from_texts([[t] for _, t in ex["queries"]])
    res = db.align_batch(ex["params"], batch)
    store = vo.Store(contigs, 0.1)
    vo.accumulate(store, res, synth.unpack_reads(batch), names)
    per_query = [int(res["comp_choice_off"][res["q_comp_off"][q] + 1] - res["comp_choice_off"][res["q_comp_off"][q]]) for q in range(len(names))]
    return dict(source="oracle-generated (see make_examples_expected.py); inputs: examples/reference.fasta, examples/queries.fasta, flags of examples/test.sh:14",
                contig_order=[n for n, _ in contigs], choices_per_query=per_query,
                sam_body=sam_oracle.format_sam(res, batch, names, [n for n, _ in contigs]),
                vcf_body=vo.vcf_body(store), mutations_body=vo.mutations_body(store, vo.Filter.default()))


if __name__ == "__main__":
    out = build()
    with open(os.path.join(HERE, "examples_expected.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(out["sam_body"])
    print(out["mutations_body"])
    print(out["vcf_body"][:600])
