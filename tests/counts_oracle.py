"""ORACLE (test infrastructure only): numpy/pure-Python restatement of the reference's reference-base depth
accumulation — QV/MatchDatabase.java:34-59 (weight = 1/numChoices as a float), QV/Alignments.java:89-156
(per aligned base; isNearQueryEnd), QV/WeightedAlignment.java:19-28, QV/QueryAlignment.java:97-120,203-214
(numAlignmentsCoveringIndexB via minOverlap/maxOverlap), QV/DirectionalAlignments.java:20-28 ((int)(weight*100)
added when the unambiguous query base equals the reference base).  Pure-Python loops: small cases only.
Pinned by T/MatchDatabase_Test.java:37-69 (tests/test_counts_oracle.py)."""
import numpy as np

F = np.float32
COMP = np.array([((c & 8) >> 3) | ((c & 4) >> 1) | ((c & 2) << 1) | ((c & 1) << 3) for c in range(16)], dtype=np.uint8)


def _is_ambiguous(code):
    return code not in (0, 1, 2, 4, 8)


def depth_planes(contig_codes, reads, results, end_fraction):
    """contig_codes: list of uint8 code arrays (forward contigs, database order); reads: list (per query) of lists of uint8
    code arrays (mates as read); results: dict of the flat xm_results arrays.  Returns one int32 array [2][2][len] per contig."""
    planes = [np.zeros((2, 2, len(c)), dtype=np.int32) for c in contig_codes]
    r = results
    for q in range(len(reads)):
        if r["q_status"][q] != 0:
            continue
        comps = range(r["q_comp_off"][q], r["q_comp_off"][q + 1])
        n_comp = len(comps)
        for ci, c in enumerate(comps):
            k0, k1 = r["comp_choice_off"][c], r["comp_choice_off"][c + 1]
            if k1 - k0 < 1:
                continue
            weight = F(1.0) / F(k1 - k0)
            for k in range(k0, k1):
                sas = range(r["choice_sa_off"][k], r["choice_sa_off"][k + 1])
                n_sa = len(sas)
                min_ov = max_ov = -1
                spans = []
                for s in sas:
                    bl = r["blocks"][4 * r["sa_block_off"][s]:4 * r["sa_block_off"][s + 1]].reshape(-1, 4)
                    mn, mx = int(bl[0, 1]), int(bl[-1, 1] + bl[-1, 3])
                    if min_ov < 0 or mn >= min_ov:
                        min_ov = mn
                    if max_ov < 0 or mx <= max_ov:
                        max_ov = mx
                    spans.append(bl)
                for si, s in enumerate(sas):
                    mate = ci if n_comp == 2 else si
                    codes = reads[q][mate]
                    rev = int(r["sa_reversed"][s])
                    if rev:
                        codes = COMP[codes[::-1]]
                    contig = int(r["sa_contig"][s])
                    refc = contig_codes[contig]
                    bl = spans[si]
                    first_start_a, last_end_a = int(bl[0, 0]), int(bl[-1, 0] + bl[-1, 2])
                    limit = float(len(codes)) * float(end_fraction)
                    for a0, b0, al, blen in bl.tolist():
                        if al != blen:
                            continue
                        for i in range(al):
                            qa, rb = a0 + i, b0 + i
                            code = int(codes[qa])
                            if _is_ambiguous(code) or code != int(refc[rb]):
                                continue
                            num = n_sa
                            if n_sa >= 2 and (rb < min_ov or rb >= max_ov):
                                num = 1
                            pos_w = F(1.0) / F(num) if num != 0 else F(0)
                            wgt = F(weight * pos_w)
                            dist = min(qa - first_start_a, last_end_a - qa - 1)
                            region = 1 if dist < limit else 0
                            planes[contig][region, rev, rb] += int(F(wgt * F(100)))
    return planes
