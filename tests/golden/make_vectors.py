#!/usr/bin/env python3
"""Transcribes the reference's own JUnit known-answer tests for the alignment hot path into
language-neutral fixtures (tests/golden/junit_vectors.json).

The reference (mathjeff/Mapper @ ae7f346a) is Java and cannot run in the build container (no JVM), so its
unit tests are the only ground truth available.  Every case below cites the test it transcribes
(T/ = src/test/java/).  Inputs are built with the same string arithmetic the Java test uses (including Java
integer division) so that a transcription slip would show up as a structural difference, not a typo in a
300-character literal.

Run:  python tests/golden/make_vectors.py   (rewrites junit_vectors.json; deterministic)
"""
import json
import os

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "R": "Y", "Y": "R", "M": "K", "K": "M",
        "S": "S", "W": "W", "H": "D", "D": "H", "B": "V", "V": "B", "-": "-"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def params(mut=1, ins_start=1.5, ins_ext=0.6, del_start=1.5, del_ext=0.5, rate=0.2, amb=None, unal=None,
           span=0.0, max_matches=2147483647):
    """T/AlignerWorker_Test.java:788-799 makeParameters() is the default."""
    if amb is None:
        amb = rate
    if unal is None:
        unal = amb
    return dict(mutation=mut, ins_start=ins_start, ins_ext=ins_ext, del_start=del_start, del_ext=del_ext,
                max_error_rate=rate, ambiguity=amb, unaligned=unal, max_penalty_span=span,
                max_num_matches=max_matches)


ROUNDING = dict(mut=6, ins_start=9, ins_ext=5, del_start=6, del_ext=5, rate=1)  # T/AlignerWorker_Test.java:259-268

api_cases = []  # every case goes through Api.alignOnce(query, referenceText, parameters)


def api(name, cite, reference, seqs, p=None, expected_inner=0.0, per_penalty=1.0, expect=None):
    api_cases.append(dict(name=name, cite=cite, reference=reference, seqs=seqs, params=p or params(),
                          expected_inner=float(expected_inner), per_penalty=float(per_penalty), expect=expect or {}))


# --- T/AlignerWorker_Test.java ---
api("testIndelNotDuplicated", "T/AlignerWorker_Test.java:11-16",
    "TTAAACAGATCACCTCGCTGAGCGGGT", ["TTAAACAGATCACCCGCTGAGCGGGT"], expect=dict(count=1))

api("testPartialAmbiguity", "T/AlignerWorker_Test.java:19-31",
    "AACAGGCGGT" + "AACARGCGGT" + "AACARRCGGT", ["AACAAGCGGT"],
    expect=dict(count=1, aligned_b0="AACARGCGGT"))

_ref = "AAAAAAAAAAACGGAAAGAAATAACTTAAACGAACTAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAACGGAAAGAAATAAA"
api("testPairedEndQueries.query1", "T/AlignerWorker_Test.java:712-725", _ref, ["CGGAAAGAAA"], expect=dict(count=2))
for _rev, _n in ((True, 1), (False, 0)):
    _s2 = "CTTAAACGAACT"
    if _rev:
        _s2 = rc(_s2)
    api("testPairedEndQueries.query2.rev%s" % _rev, "T/AlignerWorker_Test.java:727-730", _ref, [_s2], expect=dict(count=1))
    api("testPairedEndQueries.combined.rev%s" % _rev, "T/AlignerWorker_Test.java:732-742", _ref, ["CGGAAAGAAA", _s2],
        expected_inner=3, per_penalty=1, expect=dict(count=_n))

_ident = "GGGGTCAC"
_q = _ident + "AAAA"
api("testHashblockAlsoMatchingNearEndOfContig", "T/AlignerWorker_Test.java:41-49",
    _ident + "CAAA" + "TCTCGGAGAGCTCGA" + _q + "T", [_q], expect=dict(count=1, aligned_b0=_q))

api("testFirstHashblockMultipleGoodMatches", "T/AlignerWorker_Test.java:52-61",
    "AACGATTTGG" + "AACGATCGCG" + "G", ["AACGATCGGG"], expect=dict(count=1, aligned_b0="AACGATCGCG"))

_q1p, _q1m, _ov, _ovm, _q2s = "AACGAGTG", "AAGGACAG", "AACGACGGTT", "AACGAGCGTT", "AAAGACCC"
api("testOverlappingPairedEndQueriesFewerMutationsOverlappingBothQueries", "T/AlignerWorker_Test.java:64-98",
    (_q1m + _ov + _q2s) + (_q1p + _ovm + _q2s), [_q1p + _ov, rc(_ov + _q2s)],
    expected_inner=0, per_penalty=1000000, expect=dict(count=1, aligned_b0=_q1p + _ovm))

_q1 = "ACGTGAACCGGTTAAACCC"
_sep = "ACAGTTGGCGAGCGC"
api("testOverlappingPairedEndQueriesBetterThanSurprisingOffset", "T/AlignerWorker_Test.java:101-144",
    _q1 + _sep + _q1 + "C", [_q1, rc(_q1)], expected_inner=0, per_penalty=len(_sep) // 2,
    expect=dict(count=2, start_b=[[0, 0], [34, 34]]))

_prefix, _shared, _sharedm, _suffix = "ACGTACGTCC", "AACCGGTTGG", "AACCTGTTGG", "AAACCCGGGTTT"
_cand = _prefix + _sharedm + _suffix
api("testOverlappingPairedEndQueriesMultipleMatches", "T/AlignerWorker_Test.java:147-172",
    "GGGG" + _cand + _cand + "TTTT", [_prefix + _shared, rc(_shared + _suffix)],
    expected_inner=0, per_penalty=len(_cand), expect=dict(count=2))

_shared = "AACCGGTTCACTCGGGACACACACC" + "ACGTCGTATTGTGCGCCGTTACAAA" + "GTTTGTTTAGAGCCCCTTTTAGCGA"
_sharedm = "AACTGGTTCACTCGGGACACACACC" + "ACGTCGTAATGTGCGCCGTTACAAA" + "GTTTGTTTAGAGCCCCTCTTAGCGA"
_cand = _sharedm
api("testMultipleCandidateMatches", "T/AlignerWorker_Test.java:175-201",
    "GGGG" + _cand + "AAAA" + _cand + "TTTT", [_shared, rc(_shared)],
    expected_inner=-1 * len(_cand), per_penalty=len(_cand) // 4, expect=dict(count=2))

_shared = "GACATTGGCAAAGTCAACAAAGCGGAAATCAAGGAAGCCATGGACGGCGTATTGAAGAAGATGCAGGGCTTTGACTTTACCAAATTCAAGGAAGAACTTGGTAAGAGAGGTTTTAAAGTCCGGGAAGCCAGGGCAAGCACCGGGAAACTC"
_cand = "T" + _shared
api("testMultipleCandidateMatches2", "T/AlignerWorker_Test.java:204-239",
    "C" + _cand + "" + _cand + "TTTT", ["G" + _shared, rc(_shared)],
    p=params(mut=6, ins_start=9, ins_ext=5.4, del_start=9, del_ext=4.5, rate=1.2),
    expected_inner=-1 * len(_cand), per_penalty=len(_cand) // 4 // 6, expect=dict(count=2))


def rounding(name, cite, q1, q2fwd, cand):
    api(name, cite, "ACGT" + cand + cand + "ACGT", [q1, rc(q2fwd)], p=params(**ROUNDING),
        expected_inner=-1 * len(cand), per_penalty=len(cand) // 4 // 6, expect=dict(count=2))


_prefix = "AAACCCGGGTTTAAAACCCCGGGGTTTTAAAAACCCCCGGGGG"
_shared = "GACATTGGCAAAGTCAACAAAGCGGAAATCAAGGAAGCCATGGACGGGGTATTGAAGAAGATGCAGGGCTTTGACTTTACCAAATTCAAGGAAGAACTTGGTAAGAG"
_sharedm = "GACATTGGCAAAGTCAACAAAGCGGAAATCAAGGAAGCCATGGACGGCGTATTGAAGAAGATGCAGGGCTTTGACTTTACCAAATTCAAGGAAGAACTTGGTAAGAG"
_suffix = "AGGTTTTAAAGTCCGGGAAGCCAGGGCAAGCACCGGGAAACTC"
rounding("testPairedEndQueriesRoundingError", "T/AlignerWorker_Test.java:242-278",
         _prefix + _sharedm, _shared + _suffix, _prefix + _shared + _suffix)

_prefix = "ATCCTTGATTTTCCCTTTAAGGGCGTTTATAATCCACCCTTTCGGATTGTTCTTTTCTCGTGATTTTCCGTTTAGGAGAGCCAGTTCTCCGATAAGGTCGGTTATCTTTTCTTGTGCCGTTATGAATGTCTCTTTGTTCCGGTTTAT"
_shared = "CTC"
_suffix = "TTCCGATGTGAAGCCGCAGGAATAACGGAGGTACTCGTACACATGGCTGTCTATCTGATATCGTGCTGTAACCTTTGCTTGCAATTCTTTCCCTTCCAGTTCTTCATCTCTGAACTGTGGGTGATAGACCGGGTAGAACCTAAACC"
_suffixm = "TTCCGATGTGAAGCCGCAGGAATAACGGAGGTACTCGTACACATGGCTGTCTATATGATATCGTGCTGTAACCTTTGCTTGCAATTCTTTCCCTTCCAGTTCTTCATCTCTGAACTGTGGGTGATAGACCGGGTAGAACCTAAACC"
rounding("testPairedEndQueriesRoundingError2", "T/AlignerWorker_Test.java:281-317",
         _prefix + _shared, _shared + _suffixm, _prefix + _shared + _suffix)

_prefix = "GAACTGGAAGGGAAAGAAT"
_shared = "TGCAAGCAAAGGTTACAGCACGATATCAGATAGACAGCCATGTGTACGAGTACCTCCGTTATTCCTGCGGCTTCACATCGGAAGAGATAAACCGGAACAAAGAGACATTCATAACGGAACAAGAAAAGATA"
_sharedm = "TGCAAGCAAAGGTTACAGCACGATATCAGATAGACAGCCATGTGTACGAGTACCTCCGTTATTCCTGCGGCTTCACATCGGAAGAGATAAACCGGAACAAAGAGACATTCATAACGGCACAAGAAAAGATA"
_suffix = "ACCGACCTTATCGGAGA"
rounding("testPairedEndQueriesRoundingError3", "T/AlignerWorker_Test.java:320-356",
         _prefix + _sharedm, _shared + _suffix, _prefix + _shared + _suffix)

_prefix = "GAACAAGGCACATGACGGTCTGGAAAACAATCCGGGAAAAGACGGCAAACT"
_prefixm = "GAACAAGGCACATGACGGTCTGGAAAACAATCCAGGAAAAGACGGCAAACT"
_shared = "GTTTTCAGACAAACACCCCTACATTACTGAAGCGCATCCGGGAGCAAAAAAAGCCGTGGACGCACTGACCAGGCGCATCAACGAAATGATAGCCGAAAT"
_suffix = "GCCGGACAACCTGACGCTGGAGGAAAAAACCGACATCGCCCGCAACAATCT"
_suffixm = "GTCGGACAACCTGACGCTGGAGGAAAAAACCGACATCGCCCGCAACAATCT"
rounding("testPairedEndQueriesRoundingError4", "T/AlignerWorker_Test.java:359-398",
         _prefixm + _shared, _shared + _suffixm, _prefix + _shared + _suffix)

_prefix = "TCTTTGTAGGGTGAAAGAGAAACCCATAAACGGGGATAGATTGAATGCTGGGAAGCATAAACAATC"
_shared = "GGGGTAAGGTTAGCGAACCTTGCCTTTCATCCCCCATTATAACTTTACATAGAGGAACTTTATCTATCCCCCCCCGCCCCCAAA"
_sharedm = "GGGGTAAGGTTAGCGTACCTTGCCTTTGATCCCCCATTATAACTTTACATAGAGGAACTTTATCTATCCCCCCCCGCCCCCAAA"
_suffix = "GGGGGAGCGACCAAACGGCAGCTTCACTCAATGGAGTGTTACAGTTCATCAAAACCAAGTGATAAC"
rounding("testPairedEndQueriesRoundingError5", "T/AlignerWorker_Test.java:401-438",
         _prefix + _shared, _sharedm + _suffix, _prefix + _shared + _suffix)

_prefix = "CAATAGGGAGATAACAGCACAAAGGATTGAGTAGAACGAAATTCGTTTGTCCACATAACCGCCGTTTTTCAT"
_suffixm = "TGTACCTTTCGGGCTGTTGCGTCCTCTATGCGCTTCGTATAGACTTCAACACGCTTTAGTTCTTGATACACC"
_suffix = "TGTACCTTTCGGGCTGTTGCGTCCTCTATGCGCTTCGTATAGACTTCAACACGCTTTAGTTCTTGATACACC"
_sharedm = "TCTGTACCCCTGCCGTTCAAAGTCCGCCAACACGTTTTTAGGCGATTTTCGGCACTTTCTAGGCTTTTCCCGTCTATT"
_shared = "TCTGTACCCCTGCCGTTCAAAGTCCGCCAACACGTTTTTTAGGCGATTTTCGGCACTTTCAAGGCTTTTCCCGTCTATT"
rounding("testPairedEndQueriesRoundingError6", "T/AlignerWorker_Test.java:441-481",
         _prefix + _sharedm, _sharedm + _suffixm, _prefix + _shared + _suffix)

_shared = "CTTCCATATCTGTTTGCTTTTAAATTCAGCACAAAGATAGCTATATTTCAATAAAATACAAACATTTTGTACACAAACGTGTACACGCCATAAAAACCCGTTTCCAATCCTACCGCCCGTTGGTTGGTTTTGCTTTGCTCTTTTTCCC"
_sharedm = "ATGCTTCCATATCTGTTTGCTTTTAAATTCAGCACAAAGATAGCTATATTTCAATAAAATACAAACATTTTGTACACAAACGTGTACACGCCATAAAAACCCGTTTCCAATCCTACCGCCCGTTGGTTGGTTTTGCTTTGCTCTTTTTCCCT"
_cand = _sharedm
api("testPairedEndQueriesOverlappingIndel", "T/AlignerWorker_Test.java:484-521",
    "ACGT" + _cand + "AACCGGTT" + _cand + "ACGT", [_shared + "CT", rc("AG" + _shared)],
    p=params(mut=6, ins_start=3, ins_ext=2, del_start=3, del_ext=2, rate=1),
    expected_inner=-1 * len(_cand), per_penalty=len(_cand) // 4 // 6, expect=dict(count=2))

_prefix = "TCTCGGCTGGCGGCAAGAGAAGAGAACACCTCGTGCAT"
_shared = "AGGCTCGCCGTTCTCTAACCAGTAAACACAATATTCGACCATAACAGTTTTATCATTTATCGTTGTAATGCCCCTCTACCTCCAAGATGTAGACCTCTACCACTTCCTCGTA"
_sharedm = "AGGCTCGCCGTTCTCTAACCAGTAAACACAATATTCGACCATAACAGTTTTATCATTTATCGTTGTAATGCCCCCTCTACCTCCAAGATGTAGACCTCTACCACTTCCTCGTA"
_suffix = "AATGTCATAGATTATCCGGTCATGGGCGGTAATGTGT"
_cand = _prefix + _shared + _suffix
api("testPairedEndQueriesOverlappingInsertion", "T/AlignerWorker_Test.java:524-562",
    "ACGT" + _cand + "ACGT" + _cand + "ACGT", [_prefix + _sharedm, rc(_sharedm + _suffix)],
    p=params(rate=0.05), expected_inner=-1 * len(_shared), per_penalty=0.5, expect=dict(count=2))

_prefix, _prefixm = "AACCGGTT", "AACCGG"
_shared = "GACATTGGCAAAGTCAACAAAGCGGAAATCAAGGAAGCCATGGACGGCGTATTGAAGAAGATGCAGGGCTTTGACTTTACCAAATTCAAGGAAGAACTTGGTAAGAGAGGTTTTAAAGTCCGGGAAGCCAGGGCAAGCACCGGGAAACTC"
_suffix, _suffixm = "AACCGGTT", "CCGGTT"
_cand = _prefixm + _shared + _suffixm
api("testPairedEndQueriesWithIndelsNextToOverlap", "T/AlignerWorker_Test.java:565-599",
    "ACGT" + _cand + "ACGT" + _cand + "ACGT", [_prefix + _shared, rc(_shared + _suffix)],
    p=params(rate=0.05), expected_inner=-1 * len(_cand), per_penalty=1, expect=dict(count=2))

_prefix = "ACCGTAACAACCTCGCAGCGTCTTTCACCAAAGCTGACAATGGCGAGCAGGTACTAATTCGCA"
_suffix = "GAAAAACGAGATTTACGCTTTGGTAAAAGTTGGTCGTGAAGATTTGATGATAACCCCGGAGCTGCAAGCAAGGATTGACAAGGCAAG"
_match = _prefix + "G" + _suffix
api("testDeletionInMiddleOfQueryWithMultipleAlignments", "T/AlignerWorker_Test.java:602-624",
    "A" + _match + _match + "A", [_prefix + _suffix], expect=dict(count=2))

api("queryExtendingPastEndOfReference", "T/AlignerWorker_Test.java:627-642",
    "GACCGGATATTCTGGTAATGACCCTTCAATTATAGACGTGAATGGTATCCAGCCGGGAGTAGATAGTAATAGTGCTTATCCTACAGCAACTCAATTGAGTTTAGGTGTGAC",
    ["ATCCTACAGCAACTCAATTGAGTTTAGGTGTGACTCTTCGCTTCAAATAAATGAGAAACAAATTATTAAAAATATGAAAGATATGAAATATATAAAATGTC"],
    expect=dict(count=1, aligned_b0="ATCCTACAGCAACTCAATTGAGTTTAGGTGTGAC"))

api("testCustomParameters", "T/AlignerWorker_Test.java:645-672", "CGCGTACTCT", ["ACGCATCCTCTTTT"],
    p=params(mut=1, ins_start=0.8, ins_ext=1, del_start=0.8, del_ext=1, rate=0.7, amb=0.9, unal=0.9),
    expect=dict(count=1, aligned_b0="CGCGTACTCT"))

_refPrefix = "A" * 77
_qPrefix, _qPrefixM = "AACACACGGTGTTCAC", "AACCCACGGTGTTCAC"
_ins = "CACCCGCCCGCGCGCTCTCTCG"
_sharedSuffix = "AATAACCGCCGGCGGTTATTAAAACCCCGGGGTTTTAAACCCGGGTTTAACCGGTTACGT"
_refSuffix = "A" * 87
assert _refPrefix == "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"
assert _refSuffix == "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"
api("testLongCheapIndel", "T/AlignerWorker_Test.java:675-695",
    _refPrefix + _qPrefixM + _sharedSuffix + _qPrefix + _refSuffix, [_qPrefix + _ins + _sharedSuffix],
    p=params(mut=2, ins_ext=0.2, del_ext=0.2),
    expect=dict(count=1, aligned_b0=_qPrefixM + "-" * len(_ins) + _sharedSuffix))

_shared = "AACCACAC"
api("test_maxPenaltySpan_with_perfectAlignment", "T/AlignerWorker_Test.java:698-710",
    _shared + "AAAA" + _shared + "AAGA", [_shared + "AAAA"], p=params(span=1), expect=dict(count=2))

# --- T/SamWriter_Test.java (hot path -> exact SAM body). DuplicationDetector(db, 1, 2, 2, 1) (:144) ---
sam_cases = []


def sam(name, cite, reference, seqs, names, expected, expected_inner=0.0, per_penalty=1.0):
    sam_cases.append(dict(name=name, cite=cite, reference=reference, ref_name="ref", seqs=seqs, names=names, params=params(),
                          expected_inner=float(expected_inner), per_penalty=float(per_penalty), expected_sam=expected,
                          dup=dict(min_len=1, max_len=2, min_copies=2, window=1)))


sam("simpleTest", "T/SamWriter_Test.java:18-30", "ACGTAAAAACCGTAAA", ["ACGTA"], ["query"],
    "query\t0\tref\t1\t255\t5M\t*\t0\t5\tACGTA\t*\tAS:f:0.0\n")
sam("pairedEndAlignment", "T/SamWriter_Test.java:33-46", "AACCGGTTATAAAAAAAAAAACGTACGTATAAAAAAAAAA",
    ["AACCGGTTAT", "ATACGTACGT"], ["one", "two"],
    "one\t99\tref\t1\t255\t10M\tref\t21\t10\tAACCGGTTAT\t*\tcs:f:0.0\tAS:f:0.0\n"
    "two\t147\tref\t21\t255\t10M\tref\t1\t10\tACGTACGTAT\t*\tcs:f:0.0\tAS:f:0.0\n", expected_inner=1, per_penalty=100)
sam("oneReadWithMultipleAlignments", "T/SamWriter_Test.java:49-61", "ACGTAAAAACGTAAAA", ["ACGTA"], ["query"],
    "query\t0\tref\t1\t255\t5M\t*\t0\t5\tACGTA\t*\tAS:f:0.0\n"
    "query\t0\tref\t9\t255\t5M\t*\t0\t5\tACGTA\t*\tAS:f:0.0\n")
sam("pairedEndReadWithMultipleAlignments", "T/SamWriter_Test.java:64-79", "ACGTAAAACCCCCTTTTACGTAAAACCCCC",
    ["ACGTA", "GGGGG"], ["one", "two"],
    "one\t99\tref\t18\t255\t5M\tref\t26\t5\tACGTA\t*\tcs:f:0.0\tAS:f:0.0\n"
    "two\t147\tref\t26\t255\t5M\tref\t18\t5\tCCCCC\t*\tcs:f:0.0\tAS:f:0.0\n"
    "one\t99\tref\t1\t255\t5M\tref\t9\t5\tACGTA\t*\tcs:f:0.0\tAS:f:0.0\n"
    "two\t147\tref\t9\t255\t5M\tref\t1\t5\tCCCCC\t*\tcs:f:0.0\tAS:f:0.0\n", expected_inner=1, per_penalty=5)
sam("pairedEndAlignmentOnlyOneSequenceAligned", "T/SamWriter_Test.java:82-94", "AACCGGTTATAAAAAAAAAAACGTACGTATAAAAAAAAAA",
    ["AACCGGTTAT", "CCCCCCCCCC"], ["one", "two"],
    "one\t73\tref\t1\t255\t10M\t*\t0\t10\tAACCGGTTAT\t*\tcs:f:0.0\tAS:f:0.0\n", expected_inner=1, per_penalty=100)

# --- T/PathAligner_Test.java (stage 8 alone, exact doubles) ---
PA = dict(mut=1, ins_start=2, ins_ext=0.5, del_start=2, del_ext=0.5, rate=0.1, amb=0.1)  # :76-87
path_cases = [
    dict(name="testQueryEndingWithMismatchAndExtension", cite="T/PathAligner_Test.java:11-15", a="AACCGGTT", b="AAT",
         aligned_a="AAC", aligned_b="AAT", penalty=1.5, params=params(**dict(PA, rate=1))),
    dict(name="testQueryStartingWithShortExtension", cite="T/PathAligner_Test.java:18-26", a="AAACCGGTTACGTACGTACGT",
         b="AACCGGTTACGTTACGTACGT", aligned_a="AACCGGTTACG-TACGTACGT", aligned_b="AACCGGTTACGTTACGTACGT", penalty=2.6,
         params=params(**dict(PA, rate=1))),
    dict(name="testMaxPenaltyHigherThanExtensionPenalty", cite="T/PathAligner_Test.java:29-39",
         a="AACACACGGTGTTCACCACCCGCCCGCGCGCT", b="AACCCACGGTGTTCACAATAACCGCCGGCGGT",
         aligned_a="AACACACGGTGTTCACCACCCGCCCGCGCGCT", aligned_b="AACCCACGGTGTTCACAATAACCGCCGGCGGT", penalty=10,
         params=params(**dict(PA, rate=1, amb=1, unal=1))),
]

# --- T/HashBlockAligner_Test.java (HashBlock_Aligner -> StraightAligner -> PathAligner), tolerance 1e-6 (:76) ---
HBA = dict(mut=1, ins_start=1.5, ins_ext=0.6, del_start=1.5, del_ext=0.5, rate=0.1, amb=0.1, max_matches=1)  # :84-96
_q = "GAGTGTCAATGACTGTTCGGCAACGGACATACTCCCGAACAGTCATTGACACTCCGTCCCACTCACGGAGAAGAGATTCTGCTGCAACCGGGCATCAACT"
_q2 = "CACGCACAATGGCATGACAGCCAACAACAAAAGTAAAAAAATCGATTTTGTTCGCATGGTAGTATTAATAGGTTTATTGATGAAGCAAAGTGTGTCTCTTAAAGAAAT"
_r3 = "TTTGATTCCTGTCTGATTCCCG"
hashblock_cases = [
    dict(name="testQueryWithLongInsertion", cite="T/HashBlockAligner_Test.java:10-17", a=_q,
         b="AAAAAAAAACAGCGCAAAGAGCTGTTCGGCAACGGACATACTCCCGAATAGTCCTTGACACTCCGTCCCACTCACGGAGAAGAGATGCTGCTGCAACCGGGCATCAACTAAAAAAAAA",
         aligned_a=_q,
         aligned_b="GAG---------CTGTTCGGCAACGGACATACTCCCGAATAGTCCTTGACACTCCGTCCCACTCACGGAGAAGAGATGCTGCTGCAACCGGGCATCAACT",
         penalty=9.9, params=params(**HBA)),
    dict(name="testInsertionCoveringThreeHashblocks", cite="T/HashBlockAligner_Test.java:20-27", a=_q2,
         b="AAAAAAAAACACGCACAATGGCATGACAGCCAACAACAAAAGTAAAAAAATCGATTTTGTTCGCATGGTAGTATTAATAGGTTTATTGATGAAGCAAAGTAAAGAAATAAATCACTTTCCCGCCAAATTTAAAAAAAAA",
         aligned_a=_q2,
         aligned_b="CACGCACAATGGCATGACAGCCAACAACAAAAGTAAAAAAATCGATTTTGTTCGCATGGTAGTATTAATAGGTTTATTGATGAAGCAAAG---------TAAAGAAAT",
         penalty=6.9, params=params(**HBA)),
    dict(name="testQueryExtendingPastEndOfReference", cite="T/HashBlockAligner_Test.java:30-38",
         a="TTTGATTCCTGTCTGATTCCCGTTCAATTCCCGCCAAGGTCCCACCGAGTTTTTTGCTTAAACCCCGTTTAATTTGCGTCAAGTTCCCGTTAAACTCCCT", b=_r3,
         aligned_a=_r3, aligned_b=_r3, penalty=7.8, params=params(**dict(HBA, rate=0.09))),
    dict(name="testQueryAlignedToMiddleOfReference", cite="T/HashBlockAligner_Test.java:41-49", a="AACGT",
         b="AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAACGTAAAAAAAAAAAAAA", aligned_a="AACGT", aligned_b="AACGT", penalty=0,
         params=params(**dict(HBA, rate=0.5))),
]

# --- T/Counting_HashBlockPath_Test.java: parameters are all-zero except DeletionExtension_Penalty = 0.1 (:71-72) ---
ZERO = dict(mut=0, ins_start=0, ins_ext=0, del_start=0, del_ext=0.1, rate=0, amb=0, unal=0)
counting_cases = [
    dict(name="checkEfficientlyHandlesRepetitionInQuery", cite="T/Counting_HashBlockPath_Test.java:12-22",
         query="GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG", reference="GGGGGGGGACGTTGCAAACCGGTTATGCTGCAAATTGGCC",
         expect=dict(num_offsets=0), params=params(**ZERO)),
    dict(name="checkOneHashblockMatchSufficientNearEndOfReference", cite="T/Counting_HashBlockPath_Test.java:25-37",
         query="CCCTTAAGGACCGTGTGAGAACGAC", reference="ACGTAAGTACGAGCCGTAAGGTCCC", expect=dict(contains_offset=12),
         params=params(**ZERO)),
    dict(name="checkPoorAlignmentInsufficientEvenNearEndOfReference", cite="T/Counting_HashBlockPath_Test.java:40-54",
         query="GGACCCGG", reference="ACCCACCCACCCACCCACCC", expect=dict(num_offsets=0), params=params(**ZERO)),
]

# --- T/HashBlockPaths_Counter_Test.java: expectedInnerDistance 10, maxInnerDistance 20 (:92-94) ---
_refText = "GGGGGACGTGGGGGGAACTAAGGGG"
_mm = "GGGGGAACAGTGGGGGGAACTAAGGGGAATTGTATATAGCG"
paths_counter_cases = [
    dict(name="checkComputesDistanceCorrectly", cite="T/HashBlockPaths_Counter_Test.java:13-18", reference=_refText,
         seq1="GACGTG", seq2=rc("AACTAAG"), expect=dict(count=1, inner=5, across=18), params=params(**ZERO)),
    dict(name="checkReverseComplementAlignment", cite="T/HashBlockPaths_Counter_Test.java:21-26", reference=rc(_refText),
         seq1="GACGTG", seq2=rc("AACTAAG"), expect=dict(count=1, inner=5, across=18), params=params(**ZERO)),
    dict(name="checkOverlappingDistance", cite="T/HashBlockPaths_Counter_Test.java:29-34", reference="GGGGAACCACTGGGGG",
         seq1="GAACCACTG", seq2=rc("CCACTGGGG"), expect=dict(count=1, inner=-6, across=12), params=params(**ZERO)),
    dict(name="checkMultipleMatches", cite="T/HashBlockPaths_Counter_Test.java:37-46", reference=_mm + _mm,
         seq1="GAACAGTG", seq2=rc("AACTAAGGGGAA"), expect=dict(count=2), params=params(**ZERO)),
]

# --- T/HashBlock_Test.java: symmetry property (:12-28) ---
symmetry_cases = ["A", "C", "G", "T", "ACGTAACCGGTTACAGATCG",
                  "TGTGTATATATAGCAAGAAGTGTCCTTGTCGGACAATTCTTGCTTTTCTCGCTTTGCTCAAAAAGATTTTAAGATTACCTTTGTGGCATGGAACTAAGACGGAACGAAAAGATTACATTCCGGTGTACCGAACTTGAAAAGGACGCACTT"]

# --- T/BasepairsTest.java:9-47 ---
basepair_cases = [dict(q="A", r="C", penalty="mutation"), dict(q="A", r="N", penalty="ambiguity"),
                  dict(q="A", r="M", penalty="ambiguity/3")]


# --- T/MutationsWriter_Test.java:18-133: seven exact mutations tables (bodies; metadata lines are stripped by the test itself, :145-155) ---
# buildMutations (:167-203): reference "ref" + its reverse complement (new SequenceDatabase(reference, true)), default HashBlock_Database,
# DuplicationDetector(db, 1, 2, 2, 1), makeAlignmentParameters (:205-216) = params() above, MatchDatabase(queryEndFraction), query "query".
_MP = params()
mutations_cases = [
    dict(name="testNoMutations", cite="T/MutationsWriter_Test.java:18-29", query="ACGTA", reference="ACGTAAAAAAAAAAAA", filter={}, end_fraction=0, expected=""),
    dict(name="testOneMutation", cite="T/MutationsWriter_Test.java:31-44", query="AACGTT", reference="AACGTAAAAA", filter={}, end_fraction=0,
         expected="ref\t6\tA\tT\t1\t1\n"),
    dict(name="testConsecutiveMutations", cite="T/MutationsWriter_Test.java:46-60", query="ACGTTTAAACCGG", reference="ACGTAAAAACCGG", filter={}, end_fraction=0,
         expected="ref\t5\tA\tT\t1\t1\nref\t6\tA\tT\t1\t1\n"),
    dict(name="testInsertion", cite="T/MutationsWriter_Test.java:62-75", query="ACGGACTTACGTCGTTAACCACGA", reference="ACGCTTACGTCGTTAACCACGA", filter={}, end_fraction=0,
         expected="ref\t3\t--\tGA\t1\t1\n"),
    dict(name="testDeletion", cite="T/MutationsWriter_Test.java:77-90", query="CACGTAACCGGTTATT", reference="CACGTAAGACCGGTTATT", filter={}, end_fraction=0,
         expected="ref\t7\tAG\t--\t1\t1\n"),
    dict(name="testIgnoringMutationWithLowDepth/filtered", cite="T/MutationsWriter_Test.java:92-105", query="ACGTAACTCCGGCTC", reference="ACGTACGTCCGGCTC",
         filter=dict(minSNPTotalDepth=2), end_fraction=0, expected=""),
    dict(name="testIgnoringMutationWithLowDepth/unfiltered", cite="T/MutationsWriter_Test.java:107-112", query="ACGTAACTCCGGCTC", reference="ACGTACGTCCGGCTC",
         filter={}, end_fraction=0, expected="ref\t6\tC\tA\t1\t1\nref\t7\tG\tC\t1\t1\n"),
    dict(name="testIgnoringIndelNearQueryEnd/filtered", cite="T/MutationsWriter_Test.java:114-128", query="CCTAACGTAACTCTGGCCGCAA", reference="AGGAACCTACGTAACTCTGGCCGCAA",
         filter=dict(minIndelTotalStartDepth=1), end_fraction=0.5, expected=""),
    dict(name="testIgnoringIndelNearQueryEnd/unfiltered", cite="T/MutationsWriter_Test.java:130-133", query="CCTAACGTAACTCTGGCCGCAA", reference="AGGAACCTACGTAACTCTGGCCGCAA",
         filter={}, end_fraction=0, expected="ref\t8\t-\tA\t1\t1\n"),
]
for _c in mutations_cases:
    _c["params"] = _MP
    _c["dup"] = dict(min_len=1, max_len=2, min_copies=2, window=1)

# --- T/MatchDatabase_Test.java:13-69: hand-built alignments, every reference position must count exactly 1 ---
match_database_cases = [
    dict(name="testQueryEndingWithMismatch", cite="T/MatchDatabase_Test.java:13-35", reference="AACCACGA", seqs=["AACCACGT"],
         alignments=[dict(b_start=0, length=8)], end_fraction=0),
    dict(name="testOverlappingPairedEndQueries", cite="T/MatchDatabase_Test.java:37-69", reference="AACCACGATTAC", seqs=["AACCACGA", "CACGATTAC"],
         alignments=[dict(b_start=0, length=8), dict(b_start=3, length=9)], end_fraction=0),
]

# --- QVT/VcfWriter_Test.java:18-80: VCF bodies from hand-built single-block alignments (SAM "name1 0 contig1 1 255 <n>M") ---
vcf_cases = [
    dict(name="simpleTest", cite="QVT/VcfWriter_Test.java:18-34", reference="ACGTAAAAACGTAAAA", seq="ACGT", support=True,
         expected="contig1\t1\tA\t.\t1\t1,0\t0,0\t.\ncontig1\t2\tC\t.\t1\t1,0\t0,0\t.\ncontig1\t3\tG\t.\t1\t1,0\t0,0\t.\ncontig1\t4\tT\t.\t1\t1,0\t0,0\t.\n"),
    dict(name="mutationTest", cite="QVT/VcfWriter_Test.java:36-53", reference="ACGTAAAAA", seq="ACGTT", support=True,
         expected="contig1\t1\tA\t.\t1\t1,0\t0,0\t.\ncontig1\t2\tC\t.\t1\t1,0\t0,0\t.\ncontig1\t3\tG\t.\t1\t1,0\t0,0\t.\ncontig1\t4\tT\t.\t1\t1,0\t0,0\t.\n"
                  "contig1\t5\tA\tT\t1\t0,0;1,0\t0,0;0,0\tACGT[T]\n"),
    dict(name="mutationWithoutSupportReads", cite="QVT/VcfWriter_Test.java:55-72", reference="ACGTAAAAA", seq="ACGTT", support=False,
         expected="contig1\t1\tA\t.\t1\t1,0\t0,0\ncontig1\t2\tC\t.\t1\t1,0\t0,0\ncontig1\t3\tG\t.\t1\t1,0\t0,0\ncontig1\t4\tT\t.\t1\t1,0\t0,0\n"
                  "contig1\t5\tA\tT\t1\t0,0;1,0\t0,0;0,0\n"),
    dict(name="oneReadWithMultipleAlignments", cite="QVT/VcfWriter_Test.java:126-145", reference="ACGTAAAAACGTAAAA", seq="ACGT", support=True, starts=[0, 8],
         expected="".join("contig1\t%d\t%s\t.\t0.5\t0.5,0\t0,0\t.\n" % (p, c) for p, c in zip([1, 2, 3, 4, 9, 10, 11, 12], "ACGTACGT"))),
    dict(name="readWithThreeAlignments", cite="QVT/VcfWriter_Test.java:181-205", reference="ACGTAAAAACGTCCCCACGT", seq="ACGT", support=True, starts=[0, 8, 16],
         expected="".join("contig1\t%d\t%s\t.\t0.33\t0.33,0\t0,0\t.\n" % (p, c) for p, c in zip([1, 2, 3, 4, 9, 10, 11, 12, 17, 18, 19, 20], "ACGTACGTACGT"))),
]

# --- T/AncestryDetector_Test.java:10-91: seven exact inferred-ancestor strings.  check() (:93-125): reference + its reverse complement,
# default HashBlock_Database, DuplicationDetector(db, chooseMin, chooseMax, 3 copies, window 0), dissimilarityThreshold 0.3,
# setVerifyNoDuplicateAnalyses() ---
def _anc(name, cite, ref, answer):
    return dict(name=name, cite=cite, reference=ref, expected=answer, threshold=0.3)


_r1, _r2 = "GCCCATTAAAACTGACACGGGTTAC", "GCCCATTAAAACTGACACCGGTTAC"
_t1, _t2 = "AACGGTGGGAACGGCGGAGCGTCGC", "AACGGTGGGATCGGCGGAGCGTCGC"
_c1, _c3 = "TTATTGTTAAACCGGTACACC", "TTATTGTTAAACCTGTACACC"
_p = ["CAACCGGAGAATCTCGATGAGNNNNNNNN", "CAACCGGAGAATCTCGATTAGNNNNNNNN", "CAACCGGAGAATCTCGATGAGNNNNNNNN", "CAACCGGAGAATCTCGATTATNNNNNNNN"]
_n1 = "GGACGTACGCACGAACGACCGAGCGATGTTT"
_m1, _m2 = "AACGACGTCTGACGAGTGACGTGGACAACCGGACGGCTC", "AACGACTTCTGACAAGTGACCTGGACATCCGGACAGCTC"
_b1, _b2, _bs = "AGCGGTGGAACGGCGGAGCGTCGTCAAACCCGGGTTCTCAGTCG", "AGCGGTGGAACGGCGGAGCGTCGTCAAACCCGGGTTCTCAGTCA", "AGACATACAGAAAGAG"
ancestry_cases = [
    _anc("basicTest", "T/AncestryDetector_Test.java:10-18", _r1 + _r1 + _r2, _r1 + _r1 + "GCCCATTAAAACTGACACSGGTTAC"),
    _anc("test2", "T/AncestryDetector_Test.java:20-28", _t1 + _t1 + _t2, _t1 + _t1 + "AACGGTGGGAWCGGCGGAGCGTCGC"),
    _anc("reverseComplementTest", "T/AncestryDetector_Test.java:30-39", _c1 + rc(_c1) + _c3, _c1 + rc(_c1) + "TTATTGTTAAACCKGTACACC"),
    _anc("proceedPastTiesTest", "T/AncestryDetector_Test.java:41-54", "".join(_p), _p[0] + _p[1] + _p[2] + "CAACCGGAGAATCTCGATTAKNNNNNNNN"),
    _anc("noChangesTest", "T/AncestryDetector_Test.java:56-64", _n1 * 3, _n1 * 3),
    _anc("manyMutationsTest", "T/AncestryDetector_Test.java:66-76", _m1 + _m1 + _m2, _m1 + _m1 + "AACGACKTCTGACRAGTGACSTGGACAWCCGGACRGCTC"),
    _anc("breakSimilarSectionTest/mutatedAtEnd", "T/AncestryDetector_Test.java:78-85", _b1 * 3 + _b2 + _bs, _b1 * 3 + _b2 + _bs),
    _anc("breakSimilarSectionTest/mutatedInMiddle", "T/AncestryDetector_Test.java:87-91", _b1 + _b1 + _b2 + _b1 + _bs,
         _b1 + _b1 + "AGCGGTGGAACGGCGGAGCGTCGTCAAACCCGGGTTCTCAGTCR" + _b1 + _bs),
]

# --- examples/ (config 1): inputs only; the reference ships no expected output (examples/.gitignore) ---
examples = dict(
    cite="examples/reference.fasta, examples/queries.fasta, examples/test.sh:14",
    reference=[["contig1", "AAAACCAAAGGCTCGCGTA"], ["contig2", "ACGTAC"], ["contig3", "ACGTAACCGGTTAAACCCGGGTTTAAAACCCCGGGGTTTT"]],
    queries=[["query1-matches", "AAAACCAAAGG"], ["query2-1SNP", "AAAACCAAATG"], ["query3-matches", "ACGTAC"],
             ["query4-insertion", "AAAACCCAAAGG"], ["query5-deletion", "CCGGTTAAACCCGGTTTAAAACCCC"],
             ["query6-too-different", "ACGCGCTAAACCGAGG"]],
    # M/Mapper.java:409-453 defaults
    params=params(mut=1, ins_start=1.5, ins_ext=0.6, del_start=1.5, del_ext=0.5, rate=0.1, amb=0.1, unal=0.1, span=0.5))

out = dict(source="mathjeff/Mapper @ ae7f346a JUnit tests (transcribed; see make_vectors.py)",
           api_cases=api_cases, sam_cases=sam_cases, path_aligner_cases=path_cases, hashblock_aligner_cases=hashblock_cases,
           counting_path_cases=counting_cases, paths_counter_cases=paths_counter_cases, symmetry_cases=symmetry_cases,
           basepair_cases=basepair_cases, examples=examples,
           mutations_cases=mutations_cases, match_database_cases=match_database_cases, vcf_cases=vcf_cases, ancestry_cases=ancestry_cases)

if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "junit_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", path, "api_cases=%d" % len(api_cases))
