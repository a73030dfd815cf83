"""Shared helpers for parity tests: feeding a target (the product through its C ABI, or the host emulation harness)
from the oracle acting as the reference's Java host, and comparing flat result arrays bit for bit."""
import numpy as np

from mapper_b200 import synth

COMPARE = ["q_comp_off", "comp_choice_off", "choice_sa_off", "sa_block_off", "choice_f64", "sa_f64", "choice_inner", "sa_contig",
           "blocks", "q_status", "sa_reversed"]


def batch_from_texts(queries, expected_inner=None, per_penalty=None):
    """queries: list of [text] or [text1, text2]."""
    reads = []
    for q in queries:
        for s in q:
            reads.append(np.array([synth.LETTERS.tobytes().index(c.encode()) for c in s.upper()], dtype=np.uint8))
    packed, off, lens = synth.pack_reads(reads)
    if len(packed) == 0:
        packed = np.zeros(1, dtype=np.uint16)
    nq = len(queries)
    ei = np.zeros(nq) if expected_inner is None else np.asarray(expected_inner, dtype=np.float64)
    pp = np.ones(nq) if per_penalty is None else np.asarray(per_penalty, dtype=np.float64)
    return dict(packed=packed, seq_word_off=off, seq_len=lens, n_seqs=np.array([len(q) for q in queries], dtype=np.uint8),
                expected_inner=np.ascontiguousarray(ei, dtype=np.float64), per_penalty=np.ascontiguousarray(pp, dtype=np.float64))


def feed_reference(target, oracle_db):
    packed, lens = [], []
    for i in range(oracle_db.num_contigs()):
        _, codes = oracle_db.contig(i)
        packed.append(synth.pack_contig(codes))
        lens.append(len(codes))
    target.set_reference(packed, lens)


def feed_from_oracle(target, oracle_db, max_used, window, upload_index=True, upload_dups=True):
    """The oracle plays the Java host: its HashBlock_Database tables and duplication keys are uploaded."""
    feed_reference(target, oracle_db)
    if upload_index:
        built = oracle_db.build_through(max_used)
        for n in range(0, built + 1):
            target.set_index_length(oracle_db.table(n))
        target.finish_index(oracle_db.min_interesting(), built)
    if upload_dups:
        oracle_db.detect_duplications()
        for c in range(oracle_db.num_contigs()):
            target.set_duplications(window, oracle_db.dup_granularity(), c, oracle_db.dup_starts(c))


def assert_same_results(a, b, what=""):
    for k in COMPARE:
        x, y = a[k], b[k]
        if x.shape != y.shape or not np.array_equal(x, y):
            # locate the first differing query for a useful message
            nq = len(a["q_status"])
            msg = "%s: array %s differs (shapes %s vs %s)" % (what, k, x.shape, y.shape)
            for q in range(nq):
                if describe(a, q) != describe(b, q):
                    msg += "\nfirst differing query %d:\n  A: %s\n  B: %s" % (q, describe(a, q), describe(b, q))
                    break
            raise AssertionError(msg)


def describe(r, q):
    """Structured view of query q's alignments."""
    out = dict(status=int(r["q_status"][q]), comps=[])
    for c in range(r["q_comp_off"][q], r["q_comp_off"][q + 1]):
        choices = []
        for k in range(r["comp_choice_off"][c], r["comp_choice_off"][c + 1]):
            sas = []
            for s in range(r["choice_sa_off"][k], r["choice_sa_off"][k + 1]):
                blocks = r["blocks"][4 * r["sa_block_off"][s]:4 * r["sa_block_off"][s + 1]].reshape(-1, 4).tolist()
                sas.append(dict(contig=int(r["sa_contig"][s]), reversed=int(r["sa_reversed"][s]), penalty=float(r["sa_f64"][2 * s]).hex(),
                                aligned=float(r["sa_f64"][2 * s + 1]).hex(), blocks=blocks))
            choices.append(dict(f64=[float(v).hex() for v in r["choice_f64"][4 * k:4 * k + 4]], inner=int(r["choice_inner"][k]), sas=sas))
        out["comps"].append(choices)
    return out


def inferred_ancestor_oracle(ref, params, threads=4, window=1000):
    """--infer-ancestors as M/Mapper.java:666-692 sets it up, with the oracle playing the Java host: the original reference hashed with
    minInterestingSize = minDuplicationLength and 8 short matches, DuplicationDetector(3 copies, window 1), AncestryDetector with
    dissimilarityThreshold = MaxErrorRate / MutationPenalty; the aligner then works on the "-anc" reference (IUPAC unions), whose own
    duplication detector uses 2 copies and the 1000-base window.  ref: list of (name, codes).  Returns (anc oracle, changed positions)."""
    import xm_oracle as xo
    texts = [(n, synth.codes_to_text(s)) for n, s in ref]
    total = sum(len(s) for _, s in ref)
    min_dup = 1
    while (1 << min_dup) < total:   # SequenceDatabase.log2RoundUp(totalForwardSize), M/DuplicationDetector.java:17-31
        min_dup += 1
    db0 = xo.Oracle(texts, sort_by_length=True, min_interesting=min_dup, max_short=8, threads=threads, dup=dict(min_len=min_dup, max_len=2 * min_dup, min_copies=3, window=1))
    anc = db0.infer_ancestors(params["max_error_rate"] / params["mutation"])
    changed = sum(int((a[1] != db0.contig(i)[1]).sum()) for i, a in enumerate(anc))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in anc], sort_by_length=False, threads=threads, dup=dict(min_copies=2, window=window))
    return db, changed
