"""ORACLE - TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of QV/SamWriter.java:118-352 (one SAM line per sequence
alignment) over the flat result arrays, pinned by the five exact SAM bodies of T/SamWriter_Test.java
(tests/golden/junit_vectors.json, sam_cases).  Used to check xm_format_sam."""
import math

import numpy as np

LETTERS = "-ACMGRSVTWYHKDBN"  # QV/Basepairs.java decode
COMP = [((c & 8) >> 3) | ((c & 4) >> 1) | ((c & 2) << 1) | ((c & 1) << 3) for c in range(16)]


def java_float_str(v):
    """Float.toString for a finite float32 (shortest digits; decimal form for 1e-3 <= |v| < 1e7, else d.dddE<n>)."""
    v = np.float32(v)
    if v == 0:
        return "-0.0" if np.signbit(v) else "0.0"
    a = abs(float(v))
    if 1e-3 <= a < 1e7:
        s = np.format_float_positional(v, unique=True, trim="0")
        if s.endswith("."):
            s += "0"
        return s
    s = np.format_float_scientific(v, unique=True, trim="0", exp_digits=1)  # like 1.e-04 / 1.5e+07
    mant, exp = s.split("e")
    if mant.endswith("."):
        mant += "0"
    return "%sE%d" % (mant, int(exp))


def format_number(penalty):  # formatSequencePenalty / formatQueryPenalty / formatNumber :263-277
    score = np.float32(-1 * penalty)
    scaled = float(score) * 10000.0
    r = math.floor(scaled + 0.5)  # Math.round(double)
    return "f:" + java_float_str(np.float32(np.float32(r) / np.float32(10000)))


def unpack(batch, sid):
    off = int(batch["seq_word_off"][sid]); n = int(batch["seq_len"][sid])
    w = batch["packed"][off:off + (n + 3) // 4].astype(np.uint32)
    return np.stack([(w >> s) & 15 for s in (0, 4, 8, 12)], 1).ravel()[:n]


def format_sam(r, batch, seq_names, contig_names):
    out = []
    first_seq = np.concatenate([[0], np.cumsum(batch["n_seqs"].astype(np.int64))])
    for q in range(len(batch["n_seqs"])):
        comp0, comp1 = int(r["q_comp_off"][q]), int(r["q_comp_off"][q + 1])
        n_comp = comp1 - comp0
        having = sum(1 for c in range(comp0, comp1) if r["comp_choice_off"][c + 1] > r["comp_choice_off"][c])
        for c in range(comp0, comp1):
            sub = c - comp0
            ks = range(int(r["comp_choice_off"][c]), int(r["comp_choice_off"][c + 1]))
            min_pen = float(2147483647)
            for k in ks:
                min_pen = min(min_pen, float(r["choice_f64"][4 * k + 3]))
            for k in ks:
                pen = float(r["choice_f64"][4 * k + 3])
                has_min = (pen - min_pen) <= abs(min_pen) / 100000
                a0, a1 = int(r["choice_sa_off"][k]), int(r["choice_sa_off"][k + 1])
                n_sa = a1 - a0
                multi = n_sa > 1 or n_comp > 1
                for a in range(a0, a1):
                    i = a - a0
                    mate = sub if n_comp == 2 else i
                    sid = int(first_seq[q]) + mate
                    codes = unpack(batch, sid)
                    qlen = len(codes)
                    rev = bool(r["sa_reversed"][a])
                    other = (a0 + 1 if a == a0 else a0) if n_sa == 2 else None
                    flags = 16 if rev else 0
                    if multi:
                        flags += 1
                        if n_sa > 1:
                            flags += 2
                            if other is not None and r["sa_reversed"][other]:
                                flags += 32
                        if not (n_sa > 1 or having > 1):
                            flags += 8
                        flags += 64 if (sub + i) == 0 else 128
                    if not has_min:
                        flags += 256
                    blk = r["blocks"][4 * int(r["sa_block_off"][a]):4 * int(r["sa_block_off"][a + 1])].reshape(-1, 4)
                    cigar, consumed = "", 0
                    for a_start, _, a_len, b_len in blk.tolist():
                        if a_start != consumed:
                            cigar += "%dS" % a_start
                            consumed = a_start
                        cigar += "%dM" % a_len if a_len == b_len else ("%dI" % a_len if a_len > b_len else "%dD" % b_len)
                        consumed += a_len
                    if consumed < qlen:
                        cigar += "%dS" % (qlen - consumed)
                    if other is not None:
                        nxt = "%s\t%d" % (contig_names[int(r["sa_contig"][other])], int(r["blocks"][4 * int(r["sa_block_off"][other]) + 1]) + 1)
                    else:
                        nxt = "*\t0"
                    if rev:
                        codes = np.array([COMP[x] for x in codes[::-1]], dtype=np.int64)
                    text = "".join(LETTERS[int(x)] for x in codes)
                    tags = ("cs:" + format_number(pen) + "\t" if multi else "") + "AS:" + format_number(float(r["sa_f64"][2 * a]))
                    out.append("%s\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t%s\t*\t%s\n" % (
                        seq_names[sid], flags, contig_names[int(r["sa_contig"][a])], int(blk[0][1]) + 1, 255 if has_min else 0, cigar, nxt, qlen, text, tags))
    return "".join(out)
