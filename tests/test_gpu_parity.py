"""Parity tests proper: the CUDA library through its C ABI vs the oracle, bit for bit (integers, blocks, doubles)."""
import json
import os

import numpy as np
import pytest

import parity
import xm_oracle as xo
from mapper_b200 import capi, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "junit_vectors.json")))


def gpu_from_oracle(db, params, max_used, window, upload_index, upload_dups, threads=0):
    g = capi.XMapper(params, device=0)
    parity.feed_from_oracle(g, db, max_used, window, upload_index, upload_dups)
    if not upload_index:
        g.build_index(max_used, threads=threads)
    if not upload_dups:
        g.build_duplications(-1, -1, 2, window)
    return g


def test_junit_api_cases_on_gpu():
    """Every Api.alignOnce case of the reference's AlignerWorker_Test (host uploads its tables, as the Java host would)."""
    ran = 0
    for case in V["api_cases"]:
        db = xo.Oracle([("reference-0", case["reference"])], dup=dict(min_copies=2, window=1))
        batch = parity.batch_from_texts([case["seqs"]], [case["expected_inner"]], [case["per_penalty"]])
        g = gpu_from_oracle(db, case["params"], max(len(s) for s in case["seqs"]) + 2, 1, True, True)
        got = g.align_batch(batch, strict=True)
        want = db.align_batch(case["params"], batch)
        parity.assert_same_results(want, got, case["name"])
        n = got["comp_choice_off"][1] - got["comp_choice_off"][0] if len(got["comp_choice_off"]) == 2 else 0
        assert n == case["expect"]["count"], case["name"]
        g.close()
        ran += 1
    assert ran >= 25


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_random_reads_vs_oracle(paired):
    ref = synth.random_reference(1000000, seed=41, n_contigs=3, repeat_fraction=0.06, repeat_len=(200, 2000))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = synth.simulate_reads(contigs, 30000, 150, seed=42 + paired, sub_rate=0.015, indel_rate=0.002, paired=paired)
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)  # library builds index + duplications itself
    got = g.align_batch(batch, strict=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8)
    parity.assert_same_results(want, got, "random")
    assert got["stats"][capi.STAT["launches"]] >= 3
    g.close()


def test_edge_cases():
    """Empty batch, reads shorter than any seed, reads hanging off contig ends, a read with N (rejected loudly)."""
    ref = synth.random_reference(50000, seed=51, n_contigs=2)
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 200, 1000, False, False)
    empty = parity.batch_from_texts([])
    r = g.align_batch(empty, strict=True)
    assert len(r["q_status"]) == 0
    c0 = synth.codes_to_text(db.contig(0)[1])
    queries = [["ACGT"], [c0[:60]], [c0[-80:] + "ACGTACGTACGTAAAC"], ["TTTTGGGG" + c0[:90]], [c0[1000:1200]], [c0[2000:2100], c0[2300:2400]]]
    batch = parity.batch_from_texts(queries, [0, 0, 0, 0, 0, 300], [1, 1, 1, 1, 1, 50])
    got = g.align_batch(batch, strict=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch)
    parity.assert_same_results(want, got, "edges")
    # a read with one N aligns (status 0) exactly like the oracle's; a zero-length read is refused by both with Q_INTERNAL (-6): no
    # reference test covers it and the reference's behaviour there is an unchecked exception
    odd = parity.batch_from_texts([[c0[100:150] + "N" + c0[151:220]], [""], [c0[300:400], ""], [c0[500:580]]])
    r = g.align_batch(odd)
    want = db.align_batch(synth.DEFAULT_PARAMS, odd)
    assert r["q_status"].tolist() == [0, -6, -6, 0] and want["q_status"].tolist() == [0, -6, -6, 0]
    parity.assert_same_results(want, r, "edge statuses")
    # reads with more than 64 IUPAC-ambiguous bases (masked low-quality tails) align like the reference's: conditions are short key lists
    many_n = parity.batch_from_texts([["N" * 70 + c0[600:680]], [c0[700:740] + "N" * 66 + c0[806:850]], ["RYRYRYRYRY" * 7 + c0[900:980]]])
    parity.assert_same_results(db.align_batch(synth.DEFAULT_PARAMS, many_n), g.align_batch(many_n, strict=True), "many ambiguous bases")
    g.close()


@pytest.mark.parametrize("threads,max_entries", [(0, 0), (0, 300000), (4, 0)], ids=["device", "device-chunked", "host"])
def test_library_index_builder_on_gpu_box(threads, max_entries, monkeypatch):
    """xm_build_index: the device builder (threads=0) and the host builder it is checked against both reproduce the oracle's
    HashBlock_Database tables (M/HashBlock_Database.java:490-665, M/PackedMap.java:99-153): capacity, max count, overfull buckets,
    bucket offsets and positions.  device-chunked: XM_INDEX_MAX_ENTRIES makes the device builder work through the block lengths in
    several sorts (what it does by itself when a reference has more than 2^31 index entries or they do not fit the device); the
    alignments that read those tables in place must equal the oracle's too."""
    if max_entries:
        monkeypatch.setenv("XM_INDEX_MAX_ENTRIES", str(max_entries))
    ref = synth.random_reference(300000, seed=61, n_contigs=3, repeat_fraction=0.1, repeat_len=(100, 800))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    built = db.build_through(100)
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
    parity.feed_reference(g, db)
    g.build_index(100, threads=threads)
    mi, mb = g.index_info()
    assert mi == db.min_interesting()
    for n in range(1, min(built, mb) + 1):
        t0, t1 = db.table(n), g.get_index_length(n)
        assert t0["capacity"] == t1["capacity"] and t0["max_count"] == t1["max_count"], n
        assert np.array_equal(t0["overfull"], t1["overfull"]) and np.array_equal(t0["positions"], t1["positions"]), n
        assert np.array_equal(t0["offsets"], t1["offsets"]), n
    if max_entries:
        g.build_duplications(-1, -1, 2, 1000)
        contigs = [db.contig(i) for i in range(db.num_contigs())]
        batch = synth.simulate_reads(contigs, 3000, 100, seed=62, sub_rate=0.02, indel_rate=0.003)
        parity.assert_same_results(db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8), g.align_batch(batch, strict=True), "chunked index build")
    g.close()


@pytest.mark.parametrize("ambiguous", [False, True], ids=["plain", "anc"])
def test_duplication_table_device_scan(ambiguous):
    """xm_build_duplications (bucket scan on the device + the reference-ordered merge) == xm_build_duplications_host == the oracle's
    DuplicationDetector (M/DuplicationDetector.java:129-214, :332-436), on a reference with 2-6-copy repeat families, for window sizes 1 and 1000 and explicit length ranges."""
    ref = synth.random_reference(500000, seed=171, n_contigs=4, repeat_fraction=0.15, repeat_copies=(2, 6), repeat_len=(60, 2500))
    if ambiguous:
        db, changed = parity.inferred_ancestor_oracle(ref, synth.DEFAULT_PARAMS, threads=8)
        assert changed > 100
    else:
        db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=0)
    parity.feed_reference(g, db)
    g.build_index(150)
    db.detect_duplications()
    total = 0
    for window in (1000, 1):
        g.build_duplications(-1, -1, 2, window)
        dev = [g.get_duplications(c).copy() for c in range(db.num_contigs())]
        g.build_duplications(-1, -1, 2, window, host=True)
        for c in range(db.num_contigs()):
            assert np.array_equal(dev[c], g.get_duplications(c)), (window, c)
            if window == 1000:
                assert np.array_equal(dev[c], db.dup_starts(c)), c
            total += len(dev[c])
    assert total > 200
    for lo, hi, copies in ((12, 20, 3), (25, 64, 2)):
        g.build_duplications(lo, hi, copies, 100)
        dev = [g.get_duplications(c).copy() for c in range(db.num_contigs())]
        g.build_duplications(lo, hi, copies, 100, host=True)
        for c in range(db.num_contigs()):
            assert np.array_equal(dev[c], g.get_duplications(c)), (lo, hi, c)
    with pytest.raises(capi.XmError):
        g.build_duplications(60, 80, 2, 100)   # above 64: the device scan says so instead of switching to the host detector
    g.close()


def test_positions_past_2_32(monkeypatch):
    """Global positions wider than 32 bits (QV/SequenceDatabase.java:69-74; KAT T/PackedMap_Test.testLargeReferenceSize): with
    XM_POSITION_BIAS the reference sits 2^33 + 12345 bases into the position space.  The device builder (64-bit sort values, uint32 +
    uint8 planes), the host builder and uploaded 64-bit tables give the unbiased tables shifted by the bias; the device duplication scan
    gives the same table; alignments (single and paired) are bit-identical to the oracle's; the 32-bit entry points refuse."""
    bias = (1 << 33) + 12345
    ref = synth.random_reference(400000, seed=221, n_contigs=3, repeat_fraction=0.1, repeat_len=(100, 1500))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    plain = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)
    monkeypatch.setenv("XM_POSITION_BIAS", str(bias))
    wide = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)
    host = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False, threads=4)
    mi, mb = wide.index_info()
    assert (mi, mb) == plain.index_info()
    with pytest.raises(capi.XmError):
        wide.get_index_length(mi)
    tables = []
    for n in range(1, mb + 1):
        t0, t1, t2 = plain.get_index_length(n), wide.get_index_length(n, wide=True), host.get_index_length(n, wide=True)
        assert t0["capacity"] == t1["capacity"] and np.array_equal(t0["offsets"], t1["offsets"]) and np.array_equal(t0["overfull"], t1["overfull"]), n
        assert np.array_equal(t0["positions"].astype(np.uint64) + np.uint64(bias), t1["positions"]), n
        assert np.array_equal(t1["positions"], t2["positions"]) and np.array_equal(t1["offsets"], t2["offsets"]), n
        tables.append(t1)
    for c in range(db.num_contigs()):
        assert np.array_equal(plain.get_duplications(c), wide.get_duplications(c)), c
    up = capi.XMapper(synth.DEFAULT_PARAMS, device=0)   # the Java host's path: tables uploaded with 64-bit positions
    parity.feed_reference(up, db)
    with pytest.raises(capi.XmError):
        up.set_index_length(dict(tables[-1], positions=tables[-1]["positions"].astype(np.uint32)))
    for t in tables:
        up.set_index_length(t, wide=True)
    up.finish_index(mi, mb)
    for c in range(db.num_contigs()):
        up.set_duplications(1000, db.dup_granularity(), c, plain.get_duplications(c))
    for paired in (False, True):
        batch = synth.simulate_reads(contigs, 5000, 150, seed=222 + paired, sub_rate=0.02, indel_rate=0.003, paired=paired)
        want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8)
        for g, what in ((wide, "device-built"), (host, "host-built"), (up, "uploaded")):
            parity.assert_same_results(want, g.align_batch(batch, strict=True), "positions past 2^32, %s index, paired=%s" % (what, paired))
    for g in (plain, wide, host, up):
        g.close()


def test_long_reads_1kbp_split_shape():
    """BASELINE.json configs[4] shape: 1 kbp pieces (what --split-queries-past-size 1000 hands the aligner, M/SequenceSplitter.java:9-38)
    with 1 % substitutions + 0.5 % indels on a multi-contig reference with repeat families."""
    ref = synth.random_reference(400000, seed=71, n_contigs=3, repeat_fraction=0.05, repeat_len=(300, 3000))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = synth.simulate_reads(contigs, 1500, 1000, seed=72, sub_rate=0.01, indel_rate=0.005)
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 1000, 1000, False, False)
    got = g.align_batch(batch, strict=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8)
    parity.assert_same_results(want, got, "1 kbp reads")
    g.close()


@pytest.mark.parametrize("threads", [0, 8], ids=["device", "host"])
def test_iupac_ambiguous_reference(threads):
    """An "-anc" reference (--infer-ancestors writes IUPAC unions into the reference, M/AncestryDetector.java:323-327) produced by the
    oracle's AncestryDetector from a reference with 3-6-copy repeat families; the LIBRARY builds the index (MultiHashBlock fan-out of the
    reference, M/HashBlock_ParentRow.java:97-191 + PackedMap.add(preventDuplicates)) and the duplication table; tables and alignments
    equal the oracle's, single and paired; the ambiguity penalty decides.  threads=0: the device builder (one warp per ambiguous
    slice runs the query path's MultiHashBlock pyramid, xm_index_emit_amb_kernel + xm_index_dedupe_kernel); threads>0: the host builder."""
    ref = synth.random_reference(600000, seed=91, n_contigs=4, repeat_fraction=0.1, repeat_copies=(3, 6), repeat_len=(300, 2500))
    db, changed = parity.inferred_ancestor_oracle(ref, synth.DEFAULT_PARAMS, threads=8)
    assert changed > 200
    contigs = [(n, s) for n, s in (db.contig(i) for i in range(db.num_contigs()))]
    clean = {n: s for n, s in ref}
    sample_from = [(n, clean[n[:-4]]) for n, _ in contigs]   # reads come from the ORIGINAL reference (names: contigN-anc)
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False, threads=threads)
    built = db.build_through(150)
    for n in range(1, built + 1):
        t0, t1 = db.table(n), g.get_index_length(n)
        assert t0["capacity"] == t1["capacity"] and np.array_equal(t0["overfull"], t1["overfull"]) and np.array_equal(t0["positions"], t1["positions"]), n
    for paired in (False, True):
        batch = synth.simulate_reads(sample_from, 6000, 150, seed=92 + paired, sub_rate=0.01, indel_rate=0.002, paired=paired)
        got = g.align_batch(batch, strict=True)
        want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8)
        parity.assert_same_results(want, got, "inferred-ancestor reference paired=%s" % paired)
    g.close()


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_depth_planes_vs_counts_oracle(paired):
    """xm_counts_* (QV/MatchDatabase -> Alignments -> DirectionalAlignments reference-base depth) against tests/counts_oracle.py,
    accumulated over two batches, plus the size-independent check: total depth == 100 x matching aligned bases / mates covering."""
    import counts_oracle
    ref = synth.random_reference(120000, seed=101, n_contigs=3, repeat_fraction=0.1, repeat_len=(200, 1500), repeat_divergence=0.0)
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=4, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 120, 1000, False, False)
    g.counts_enable(0.1)
    want = [np.zeros((2, 2, len(c)), dtype=np.int32) for _, c in contigs]
    for it in range(2):
        batch = synth.simulate_reads(contigs, 1200, 120, seed=102 + 2 * it + paired, sub_rate=0.015, indel_rate=0.003, paired=paired, inner_mean=60.0, inner_sd=40.0)
        got = g.align_batch(batch, strict=True)
        planes = counts_oracle.depth_planes([c for _, c in contigs], synth.unpack_reads(batch), got, 0.1)
        for a, b in zip(want, planes):
            a += b
    total = 0
    for c in range(len(contigs)):
        have = g.counts_fetch(c)
        assert np.array_equal(have, want[c]), "contig %d" % c
        total += int(have.sum())
    assert total > 0
    g.close()


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_reads_with_ambiguous_bases(paired):
    """IUPAC-ambiguous QUERY bases (MultiHashBlocks, M/HashBlock_ParentRow.java:69-191; skipMultiblocks, M/HashBlockPath.java:130-140)."""
    from test_emu_parity import ambiguate
    ref = synth.random_reference(300000, seed=131, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 1000))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = ambiguate(synth.simulate_reads(contigs, 8000, 150, seed=132 + paired, sub_rate=0.01, indel_rate=0.002, paired=paired), 133, 0.01)
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)
    got = g.align_batch(batch, strict=True)
    want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=8)
    parity.assert_same_results(want, got, "ambiguous reads paired=%s" % paired)
    g.close()


def test_sam_bodies_on_gpu():
    """T/SamWriter_Test.java: the five exact SAM bodies, aligned AND formatted on the device (xm_format_sam)."""
    for case in V["sam_cases"]:
        db = xo.Oracle([(case["ref_name"], case["reference"])], sort_by_length=False, dup=case["dup"])
        batch = parity.batch_from_texts([case["seqs"]], [case["expected_inner"]], [case["per_penalty"]])
        g = gpu_from_oracle(db, case["params"], max(len(s) for s in case["seqs"]) + 2, 1, True, True)
        _, sam = g.align_batch_sam(batch, case["names"], [case["ref_name"]], strict=True)
        assert sam == case["expected_sam"], case["name"]
        g.close()


@pytest.mark.parametrize("paired", [False, True], ids=["single", "paired"])
def test_sam_random_reads_vs_oracle(paired):
    """Device SAM text == the Python restatement of QV/SamWriter.java (pinned by the JUnit bodies) on reads with substitutions,
    indels, soft clips at contig ends, ambiguous bases, unaligned mates and multiple choices."""
    import sam_oracle
    from test_emu_parity import ambiguate
    ref = synth.random_reference(120000, seed=141, n_contigs=3, repeat_fraction=0.1, repeat_len=(150, 600))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = ambiguate(synth.simulate_reads(contigs, 4000, 100, seed=142 + paired, sub_rate=0.02, indel_rate=0.004, paired=paired, inner_mean=80.0, inner_sd=25.0), 143, 0.003)
    assert batch["seq_len"].min() > 0
    n_seqs = int(batch["n_seqs"].astype(np.int64).sum())
    names = ["read%d/%d" % (i // (2 if paired else 1), i % 2 + 1) for i in range(n_seqs)]
    cnames = [db.contig(i)[0] for i in range(db.num_contigs())]
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 100, 1000, False, False)
    got, sam = g.align_batch_sam(batch, names, cnames, strict=True)
    want = sam_oracle.format_sam(got, batch, names, cnames)
    assert len(sam) > 100000
    if sam != want:
        a, b = sam.split("\n"), want.split("\n")
        for x, y in zip(a, b):
            assert x == y
    assert sam == want
    g.close()


@pytest.mark.parametrize("name", ["no-gapmers", "cheap-indels", "loose-error-rate", "no-span-one-match", "costly-mutation"])
def test_parameter_variants(name):
    """Non-default AlignmentParameters and --no-gapmers through the C ABI (index and duplication table built by the library)."""
    from test_emu_parity import variant_params
    p = variant_params(name)
    ref = synth.random_reference(200000, seed=151, n_contigs=2, repeat_fraction=0.08, repeat_len=(150, 800))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, gapmers=bool(p.get("enable_gapmers", 1)), dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    g = gpu_from_oracle(db, p, 120, 1000, False, False)
    for paired in (False, True):
        batch = synth.simulate_reads(contigs, 4000, 120, seed=153 + paired, sub_rate=0.02, indel_rate=0.004, paired=paired, inner_mean=150.0, inner_sd=20.0)
        got = g.align_batch(batch, strict=True)
        want = db.align_batch(p, batch, threads=8)
        parity.assert_same_results(want, got, "%s paired=%s" % (name, paired))
    g.close()


def test_consecutive_batches_and_result_lifetime():
    """One handle, many xm_align_batch calls of different sizes (0, 1, odd, larger than the previous), results of earlier batches kept
    alive (zero-copy views into the pinned slabs) while later ones run; xm_format_sam refuses results whose device copy is gone."""
    import ctypes as C
    ref = synth.random_reference(150000, seed=161, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 600))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)
    kept = []
    for i, n in enumerate([1, 0, 777, 5000, 33, 12000, 2]):
        if n == 0:
            batch = parity.batch_from_texts([])
        else:
            batch = synth.simulate_reads(contigs, n, 150, seed=170 + i, sub_rate=0.015, indel_rate=0.002, paired=(i % 2 == 1))
        got = g.align_batch(batch, strict=True, copy=False)
        want = db.align_batch(synth.DEFAULT_PARAMS, batch, threads=4)
        kept.append((want, got))
    for want, got in kept:  # every earlier result is still intact
        parity.assert_same_results(want, {k: v for k, v in got.items() if k != "_owner"}, "kept result")
    # SAM text is produced from the device copy of the results, which lives in the batch slot that served the call until that slot
    # serves another batch (three slots, least recently used first): recent results format fine, old ones are refused loudly
    batch = synth.simulate_reads(contigs, 10, 150, seed=190)
    args = lambda b: (len(b["n_seqs"]), capi._ptr(b["packed"]), capi._ptr(b["seq_word_off"]), capi._ptr(b["seq_len"]), capi._ptr(b["n_seqs"]), capi._ptr(b["expected_inner"]), capi._ptr(b["per_penalty"]))
    rs = [C.c_void_p() for _ in range(5)]
    for r in rs:
        assert g.L.xm_align_batch(g.h, *args(batch), C.byref(r)) == 0
    names = ["r%d" % i for i in range(10)]
    with pytest.raises(capi.XmError):
        g.format_sam(rs[0], names, ["c0", "c1"])
    with pytest.raises(capi.XmError):
        g.format_sam(rs[1], names, ["c0", "c1"])
    texts = [g.format_sam(r, names, ["c0", "c1"]) for r in rs[2:]]
    assert texts[0].count("\n") >= 9 and texts[0] == texts[1] == texts[2]
    r1, r2 = rs[0], rs[1]
    for r in rs[2:]:
        g.L.xm_release_results(r)
    g.L.xm_release_results(r1); g.L.xm_release_results(r2)
    g.close()


def test_concurrent_callers_on_one_handle():
    """M/Api.java:78 ("expected to be threadsafe"), M/Mapper.java:1026-1040 (N AlignerWorkers): several host threads call xm_align_batch
    on ONE handle at the same time; every call gets exactly the results of its own batch (the copies of one call overlap the kernels of
    another, the kernels themselves run one batch after the other)."""
    import threading
    ref = synth.random_reference(300000, seed=171, n_contigs=2, repeat_fraction=0.05, repeat_len=(200, 800))
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=8, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    g = gpu_from_oracle(db, synth.DEFAULT_PARAMS, 150, 1000, False, False)
    n_threads, per_thread = 4, 6
    batches = [[synth.simulate_reads(contigs, 3000 + 500 * t + 100 * k, 150, seed=700 + 10 * t + k, sub_rate=0.015, indel_rate=0.002, paired=(k % 2 == 1))
                for k in range(per_thread)] for t in range(n_threads)]
    want = [[db.align_batch(synth.DEFAULT_PARAMS, b, threads=8) for b in bs] for bs in batches]
    got = [[None] * per_thread for _ in range(n_threads)]
    errs = []

    def run(t):
        try:
            for k in range(per_thread):
                got[t][k] = g.align_batch(batches[t][k], strict=True)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(t,)) for t in range(n_threads)]
    [x.start() for x in th]
    [x.join(timeout=300) for x in th]
    assert not errs, errs
    for t in range(n_threads):
        for k in range(per_thread):
            parity.assert_same_results(want[t][k], got[t][k], "thread %d batch %d" % (t, k))
    g.close()
