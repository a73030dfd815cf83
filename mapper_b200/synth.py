"""Seeded synthetic references and simulated reads for the BASELINE.json configs.

Everything is numpy; reads are produced directly in the QV 4-bit packed layout the C-ABI takes
(base i of a sequence at bits 4*(i&3) of 16-bit word i>>2; A=1 C=2 G=4 T=8, QV/SequenceBuilder.java:20-38).
"""
import numpy as np

CODES = np.array([1, 2, 4, 8], dtype=np.uint8)
COMP = np.zeros(16, dtype=np.uint8)
for _c in range(16):
    COMP[_c] = ((_c & 8) >> 3) | ((_c & 4) >> 1) | ((_c & 2) << 1) | ((_c & 1) << 3)
LETTERS = np.frombuffer(b"-ACMGRSVTWYHKDBN", dtype=np.uint8)


def random_reference(n_bases, seed, n_contigs=1, repeat_fraction=0.0, repeat_copies=(2, 4), repeat_len=(1000, 5000),
                     repeat_divergence=0.02):
    """Returns a list of (name, uint8 codes). Contig lengths are seeded; optional repeat families."""
    rng = np.random.default_rng(seed)
    if n_contigs == 1:
        lengths = [n_bases]
    else:
        w = rng.uniform(1.0, 10.0, size=n_contigs)
        lengths = np.maximum(1000, (w / w.sum() * n_bases).astype(np.int64)).tolist()
    contigs = []
    for i, L in enumerate(lengths):
        contigs.append(["contig%d" % (i + 1), CODES[rng.integers(0, 4, size=int(L))]])
    if repeat_fraction > 0:
        target = int(n_bases * repeat_fraction)
        placed = 0
        while placed < target:
            rl = int(rng.integers(repeat_len[0], repeat_len[1] + 1))
            copies = int(rng.integers(repeat_copies[0], repeat_copies[1] + 1))
            unit = CODES[rng.integers(0, 4, size=rl)]
            for _ in range(copies):
                c = int(rng.integers(0, len(contigs)))
                seq = contigs[c][1]
                if len(seq) <= rl + 2:
                    continue
                pos = int(rng.integers(0, len(seq) - rl))
                copy = unit.copy()
                nmut = rng.binomial(rl, rng.uniform(0, repeat_divergence))
                if nmut:
                    where = rng.integers(0, rl, size=nmut)
                    copy[where] = CODES[rng.integers(0, 4, size=nmut)]
                seq[pos:pos + rl] = copy
                placed += rl
    return [(n, s) for n, s in contigs]


def _mutate(frag, rng, sub_rate, indel_rate, max_indel=10):
    """Substitutions + small indels (geometric(0.5) length capped, insertion:deletion 1:1)."""
    L = len(frag)
    out = frag.copy()
    nsub = rng.binomial(L, sub_rate)
    if nsub:
        where = rng.integers(0, L, size=nsub)
        out[where] = CODES[(np.searchsorted(CODES, out[where]) + rng.integers(1, 4, size=nsub)) % 4]
    nind = rng.binomial(L, indel_rate)
    if nind:
        pieces = []
        prev = 0
        for p in np.sort(rng.integers(1, L - 1, size=nind)):
            if p < prev:
                continue
            ln = int(min(max_indel, rng.geometric(0.5)))
            pieces.append(out[prev:p])
            if rng.random() < 0.5:
                pieces.append(CODES[rng.integers(0, 4, size=ln)])  # insertion in the read
                prev = p
            else:
                prev = min(L, p + ln)  # deletion from the read
        pieces.append(out[prev:])
        out = np.concatenate(pieces)
    return out


def pack_reads(reads):
    """reads: list of uint8 code arrays -> (packed uint16, seq_word_off int64, seq_len int32)."""
    lens = np.array([len(r) for r in reads], dtype=np.int32)
    words = (lens.astype(np.int64) + 3) // 4
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(words, out=off[1:])
    packed = np.zeros(int(off[-1]), dtype=np.uint16)
    # vectorised when all reads have the same length
    if len(reads) and (lens == lens[0]).all():
        L = int(lens[0])
        W = int(words[0])
        m = np.zeros((len(reads), W * 4), dtype=np.uint16)
        m[:, :L] = np.stack(reads)
        m = m.reshape(len(reads), W, 4)
        packed[:] = (m[:, :, 0] | (m[:, :, 1] << 4) | (m[:, :, 2] << 8) | (m[:, :, 3] << 12)).reshape(-1)
    else:
        for i, r in enumerate(reads):
            pad = np.zeros(int(words[i]) * 4, dtype=np.uint16)
            pad[:len(r)] = r
            pad = pad.reshape(-1, 4)
            packed[off[i]:off[i + 1]] = pad[:, 0] | (pad[:, 1] << 4) | (pad[:, 2] << 8) | (pad[:, 3] << 12)
    return packed, off, lens


def unpack_reads(batch):
    """Inverse of pack_reads for a packed batch: list (per query) of lists of uint8 code arrays (mates as sent)."""
    off, lens, n_seqs = batch["seq_word_off"], batch["seq_len"], batch["n_seqs"]
    seqs = []
    for i in range(len(lens)):
        w = batch["packed"][off[i]:off[i + 1]].astype(np.uint16)
        codes = np.stack([w & 15, (w >> 4) & 15, (w >> 8) & 15, (w >> 12) & 15], axis=1).reshape(-1).astype(np.uint8)
        seqs.append(codes[:int(lens[i])])
    out, k = [], 0
    for n in n_seqs:
        out.append(seqs[k:k + int(n)])
        k += int(n)
    return out


def pack_contig(codes):
    """uint8 codes -> QV-packed uint16 words."""
    n = len(codes)
    pad = np.zeros(((n + 3) // 4) * 4, dtype=np.uint16)
    pad[:n] = codes
    pad = pad.reshape(-1, 4)
    return (pad[:, 0] | (pad[:, 1] << 4) | (pad[:, 2] << 8) | (pad[:, 3] << 12)).astype(np.uint16)


def simulate_reads(contigs, n_reads, read_len, seed, sub_rate=0.01, indel_rate=0.001, paired=False, inner_mean=300.0,
                   inner_sd=30.0, fixed_length=True):
    """Single-end or FR paired reads. Returns a batch dict for xm_align_batch plus the truth (contig, pos, strand)."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c[1]) for c in contigs], dtype=np.int64)
    prob = lens / lens.sum()
    reads = []
    truth = np.zeros((n_reads, 3), dtype=np.int64)
    cidx = rng.choice(len(contigs), size=n_reads, p=prob)
    for i in range(n_reads):
        seq = contigs[int(cidx[i])][1]
        if paired:
            inner = int(max(-100, rng.normal(inner_mean, inner_sd)))
            span = 2 * read_len + inner
        else:
            span = read_len
        span = min(span + 24, len(seq))
        pos = int(rng.integers(0, len(seq) - span + 1))
        strand = int(rng.integers(0, 2))
        truth[i] = (cidx[i], pos, strand)
        frag = seq[pos:pos + span]
        if strand:
            frag = COMP[frag[::-1]]
        if paired:
            m1 = _mutate(frag[:read_len + 12], rng, sub_rate, indel_rate)[:read_len]
            tail = frag[max(0, len(frag) - 24 - read_len - 12):len(frag) - 24]  # (a negative start would wrap around to an empty mate)
            m2 = COMP[_mutate(tail, rng, sub_rate, indel_rate)[::-1]][:read_len]
            reads.append(m1)
            reads.append(m2)
        else:
            r = _mutate(frag, rng, sub_rate, indel_rate)
            reads.append(r[:read_len] if fixed_length else r)
    packed, off, slen = pack_reads(reads)
    nq = n_reads
    return dict(packed=packed, seq_word_off=off, seq_len=slen,
                n_seqs=np.full(nq, 2 if paired else 1, dtype=np.uint8),
                expected_inner=np.full(nq, inner_mean if paired else 0.0, dtype=np.float64),
                per_penalty=np.full(nq, 50.0 if paired else 1.0, dtype=np.float64), truth=truth)


def codes_to_text(codes):
    return LETTERS[codes].tobytes().decode()


DEFAULT_PARAMS = dict(mutation=1.0, ins_start=1.5, ins_ext=0.6, del_start=1.5, del_ext=0.5, max_error_rate=0.1,
                      ambiguity=0.1, unaligned=0.1, max_penalty_span=0.5, max_num_matches=2147483647)  # M/Mapper.java:409-453


def simulate_reads_fast(contigs, n_reads, read_len, seed, sub_rate=0.01, indel_rate=0.001, paired=False, inner_mean=300.0,
                        inner_sd=30.0, per_penalty=50.0):
    """Vectorised version of simulate_reads for fixed-length reads (same error model; different random stream).
    Substitutions are applied with array ops; only reads that drew an indel event go through the Python loop."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c[1]) for c in contigs], dtype=np.int64)
    starts = np.zeros(len(contigs) + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    genome = np.concatenate([c[1] for c in contigs])
    pad = 16
    n_mates = 2 if paired else 1
    if paired:
        inner = np.maximum(-100, rng.normal(inner_mean, inner_sd, size=n_reads)).astype(np.int64)
        span = 2 * read_len + inner
    else:
        inner = np.zeros(n_reads, dtype=np.int64)
        span = np.full(n_reads, read_len, dtype=np.int64)
    cidx = rng.choice(len(contigs), size=n_reads, p=lens / lens.sum())
    room = lens[cidx] - span - pad
    assert (room > 0).all(), "contigs too short for the requested reads"
    pos = (rng.random(n_reads) * room).astype(np.int64)
    strand = rng.integers(0, 2, size=n_reads)
    W = read_len + pad
    out = np.zeros((n_reads, n_mates, read_len), dtype=np.uint8)
    cols = np.arange(W, dtype=np.int64)
    for mate in range(n_mates):
        # fragment coordinates on the forward strand of the sampled contig; reads come from the fragment ends
        frag_lo = starts[cidx] + pos
        frag_hi = frag_lo + span
        # mate 0 reads the fragment 5'->3' on its strand, mate 1 the reverse complement of the other end
        from_left = (strand == 0) == (mate == 0)
        idx = np.where(from_left[:, None], frag_lo[:, None] + cols[None, :], frag_hi[:, None] - 1 - cols[None, :])
        win = genome[idx]
        win = np.where(from_left[:, None], win, COMP[win])
        # substitutions
        n_sub = int(rng.binomial(n_reads * W, sub_rate))
        where = rng.integers(0, n_reads * W, size=n_sub)
        flat = win.reshape(-1)
        flat[where] = CODES[(np.searchsorted(CODES, flat[where]) + rng.integers(1, 4, size=n_sub)) % 4]
        win = flat.reshape(n_reads, W)
        # indels
        n_ind = rng.binomial(read_len, indel_rate, size=n_reads)
        out[:, mate, :] = win[:, :read_len]
        for i in np.nonzero(n_ind)[0]:
            row = win[i]
            pieces = []
            prev = 0
            for p in np.sort(rng.integers(1, read_len - 1, size=int(n_ind[i]))):
                if p < prev:
                    continue
                ln = int(min(10, rng.geometric(0.5)))
                pieces.append(row[prev:p])
                if rng.random() < 0.5:
                    pieces.append(CODES[rng.integers(0, 4, size=ln)])
                    prev = p
                else:
                    prev = min(W, p + ln)
            pieces.append(row[prev:])
            r = np.concatenate(pieces)
            if len(r) < read_len:
                r = np.concatenate([r, CODES[rng.integers(0, 4, size=read_len - len(r))]])
            out[i, mate, :] = r[:read_len]
    reads2d = out.reshape(n_reads * n_mates, read_len)
    Wd = (read_len + 3) // 4
    m = np.zeros((n_reads * n_mates, Wd * 4), dtype=np.uint16)
    m[:, :read_len] = reads2d
    m = m.reshape(n_reads * n_mates, Wd, 4)
    packed = (m[:, :, 0] | (m[:, :, 1] << 4) | (m[:, :, 2] << 8) | (m[:, :, 3] << 12)).reshape(-1).astype(np.uint16)
    off = np.arange(n_reads * n_mates + 1, dtype=np.int64) * Wd
    return dict(packed=packed, seq_word_off=off, seq_len=np.full(n_reads * n_mates, read_len, dtype=np.int32),
                n_seqs=np.full(n_reads, n_mates, dtype=np.uint8),
                expected_inner=np.full(n_reads, inner_mean if paired else 0.0, dtype=np.float64),
                per_penalty=np.full(n_reads, per_penalty if paired else 1.0, dtype=np.float64),
                truth=np.stack([cidx, pos, strand], axis=1))


def split_queries(reads, max_length):
    """--split-queries-past-size (M/SequenceSplitter.java:9-38): a read longer than max_length becomes
    n = (len - 1) // max_length + 1 sub-queries; piece k covers [len * k // n, len * (k + 1) // n).  Host-side, as in the reference
    (the splitter wraps the FASTA/FASTQ parser); reads: list of uint8 code arrays.  Returns (pieces, parent index per piece)."""
    pieces, parent = [], []
    for i, r in enumerate(reads):
        n = (len(r) - 1) // max_length + 1
        for k in range(n):
            pieces.append(r[len(r) * k // n:len(r) * (k + 1) // n])
            parent.append(i)
    return pieces, np.array(parent, dtype=np.int64)


def batch_from_reads(reads):
    """Single-sequence queries from a list of uint8 code arrays."""
    packed, off, slen = pack_reads(reads)
    n = len(reads)
    return dict(packed=packed, seq_word_off=off, seq_len=slen, n_seqs=np.ones(n, dtype=np.uint8),
                expected_inner=np.zeros(n, dtype=np.float64), per_penalty=np.ones(n, dtype=np.float64))
