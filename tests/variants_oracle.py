"""ORACLE - TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of QuickVariants' count accumulation and of the VCF /
mutations body formatters, over the flat result arrays of the aligner stage:

  QV/MatchDatabase.java:16-59 (weight = 1/numChoices, float), QV/WeightedAlignment.java:19-28, QV/QueryAlignment.java:97-120,203-214,
  QV/Alignments.java:89-156 (matches, insertions, deletions, isNearQueryEnd), QV/AlignmentsSection.java, QV/RegionAlignments.java,
  QV/DirectionalAlignments.java:20-160 (reference counts, alternates, insertion columns, example choice order :63-96),
  QV/Variants.java, QV/VariantsInsertions.java, QV/Variant.java, QV/AlignmentPosition.java, QV/AlignmentPosition_DirectionCounts.java,
  QV/FilteredAlignments.java, QV/MutationDetectionParameters.java, QV/MutationsFormatterWorker.java:25-147,
  QV/VcfFormatterWorker.java:27-159 (bodies only: the headers embed the command line).

Two ways to fill a Store: `accumulate()` walks alignments exactly like the Java listeners (the checker), `Store.from_device()` rebuilds
the same structure from what the CUDA library accumulated (dense reference-base planes + the sparse variant table).  Both go through
the same formatters, so equal bodies <=> equal tables.  Pinned by T/MutationsWriter_Test.java:18-133 and T/MatchDatabase_Test.java:13-69
(tests/test_variants_oracle.py).  Pure-Python loops: small cases only."""
import numpy as np

from sam_oracle import java_float_str

F = np.float32
LETTERS = "-ACMGRSVTWYHKDBN"
COMP = np.array([((c & 8) >> 3) | ((c & 4) >> 1) | ((c & 2) << 1) | ((c & 1) << 3) for c in range(16)], dtype=np.uint8)
KEYS = "ACGTN-"  # AlignmentPosition_DirectionCounts.makeKeys


def is_ambiguous_code(code):
    return code not in (0, 1, 2, 4, 8)


def is_ambiguous_char(c):
    return c not in "ACGT-"


def format_number(x):  # AlignmentPosition.formatNumber :161-167
    x = F(x)
    r = int(x)
    if r == x:
        return str(r)
    return java_float_str(x)


class Seq:
    """A query sequence as QuickVariants sees it: name, id, codes (already reverse-complemented for '-rev' views)."""
    __slots__ = ("name", "id", "codes")

    def __init__(self, name, id_, codes):
        self.name, self.id, self.codes = name, id_, codes

    def __len__(self):
        return len(self.codes)

    def text(self, a, n):
        return "".join(LETTERS[c] for c in self.codes[a:a + n])


class Variant:
    __slots__ = ("allele", "count", "ex", "ex_index")

    def __init__(self, allele):
        self.allele, self.count, self.ex, self.ex_index = allele, 0, None, 0


def better_example(v, query, qpos):  # DirectionalAlignments.betterExample :63-96
    ex = v.ex
    if ex is None:
        return True
    if len(ex) != len(query):
        return len(query) > len(ex)
    if v.ex_index != qpos:
        ideal = len(query) // 2
        d_old, d_new = abs(abs(v.ex_index) - ideal), abs(abs(qpos) - ideal)
        if d_old != d_new:
            return d_new < d_old
        return qpos < v.ex_index
    if ex.name != query.name:
        return ex.name < query.name  # existing.getName().compareTo(query.getName()) < 0
    return query.id < ex.id


class Directional:  # DirectionalAlignments
    def __init__(self, ref_codes):
        self.ref = ref_codes
        self.ref_counts = np.zeros(len(ref_codes), dtype=np.int64)
        self.alt = {}  # pos -> [list of Variant (Variants), list of lists of Variant (insertion columns)]

    def _vi(self, pos):
        v = self.alt.get(pos)
        if v is None:
            v = [[], []]
            self.alt[pos] = v
        return v

    @staticmethod
    def _get_or_create(lst, allele):
        for v in lst:
            if v.allele == allele:
                return v
        v = Variant(allele)
        lst.append(v)
        return v

    def add(self, pos, code, weight, query, qpos):  # :20-39
        if is_ambiguous_code(code):
            return
        scaled = int(F(weight) * F(100))
        if code == int(self.ref[pos]):
            self.ref_counts[pos] += scaled
            return
        v = self._get_or_create(self._vi(pos)[0], LETTERS[code])
        v.count += scaled
        if code == 0:
            qpos = -qpos
        if better_example(v, query, qpos):
            v.ex, v.ex_index = query, qpos

    def insert(self, pos, text, weight, query, qpos):  # :41-55
        rescaled = float(F(weight) * F(100))  # float * int -> float, widened to double
        cols = self._vi(pos)[1]
        for i, ch in enumerate(text):
            if is_ambiguous_char(ch):
                ch = "N"
            while len(cols) <= i:
                cols.append([])
            v = self._get_or_create(cols[i], ch)
            v.count += int(rescaled)
            if better_example(v, query, qpos + i):
                v.ex, v.ex_index = query, qpos + i


class DirCounts:  # AlignmentPosition_DirectionCounts
    __slots__ = ("ref", "ignored", "counts")

    def __init__(self):
        self.ref, self.ignored, self.counts = 0, 0, None

    def put_alt(self, key, scaled):
        if self.counts is None:
            if scaled == 0:
                return
            self.counts = [0] * 6
        self.counts[KEYS.index(key)] = scaled

    def scaled_alt(self, i):
        return 0 if self.counts is None else self.counts[i]

    def alt(self, i):
        return F(F(self.scaled_alt(i)) / F(100))

    def refc(self):
        return F(F(self.ref) / F(100))

    def ign(self):
        return F(F(self.ignored) / F(100))

    def has_alternates(self):
        return self.counts is not None and any(c != 0 for c in self.counts)


class Position:  # AlignmentPosition; containers indexed [forward][nearQueryEnd]
    def __init__(self, ref_char):
        self.ref_char = ref_char
        self.c = {(f, e): DirCounts() for f in (True, False) for e in (True, False)}
        self.sample = None  # KEYS index -> (Seq, signed index)

    def put_scaled(self, value, scaled, forward, end):
        d = self.c[(forward, end)]
        if self.ref_char == value:
            d.ref = int(scaled)
        else:
            d.put_alt(value, scaled)

    def put_sample(self, seq, index, is_deletion):  # putSampleAlternateSequence :69-92
        if is_deletion:
            alt = "-"
        else:
            code = int(seq.codes[index])
            if is_ambiguous_code(code):
                return
            alt = LETTERS[code]
        if self.sample is None:
            self.sample = [None] * 6
        self.sample[KEYS.index(alt)] = (seq, -index if is_deletion else index)

    def _sum4(self, f):  # forwardMiddle + forwardEnd + reverseMiddle + reverseEnd, float adds in that order
        return F(F(F(f(self.c[(True, False)]) + f(self.c[(True, True)])) + f(self.c[(False, False)])) + f(self.c[(False, True)]))

    def reference_count(self):
        return self._sum4(lambda d: d.refc())

    def ignored_count(self):
        return self._sum4(lambda d: d.ign())

    def alt_count_i(self, i):
        return self._sum4(lambda d: d.alt(i))

    def alt_count(self, ch):
        return self.alt_count_i(KEYS.index(ch))

    def has_alternates(self):
        return any(d.has_alternates() for d in self.c.values())

    def has_alternate(self, i):
        return any(d.alt(i) > 0 for d in self.c.values())

    def count(self):  # getCount :169-178
        t = F(self.reference_count() + self.ignored_count())
        if self.has_alternates():
            for i in range(6):
                t = F(t + self.alt_count_i(i))
        return t

    def middle_ref(self):
        return F(self.c[(True, False)].refc() + self.c[(False, False)].refc())

    def end_ref(self):
        return F(self.c[(True, True)].refc() + self.c[(False, True)].refc())

    def middle_alt_i(self, i):
        return F(self.c[(True, False)].alt(i) + self.c[(False, False)].alt(i))

    def end_alt_i(self, i):
        return F(self.c[(True, True)].alt(i) + self.c[(False, True)].alt(i))

    def middle_count(self):  # :192-201
        t = F(self.middle_ref() + F(self.c[(True, False)].ign() + self.c[(False, False)].ign()))
        if self.has_alternates():
            for i in range(6):
                t = F(t + self.middle_alt_i(i))
        return t

    def end_count(self):  # :207-215 (no ignored term)
        t = self.end_ref()
        if self.has_alternates():
            for i in range(6):
                t = F(t + self.end_alt_i(i))
        return t

    def nonzero_alternates(self):
        if not self.has_alternates():
            return []
        return [KEYS[i] for i in range(6) if self.has_alternate(i)]

    def ignore_alternate(self, ch):  # :15-27
        if ch != self.ref_char:
            i = KEYS.index(ch)
            for d in self.c.values():
                a = d.scaled_alt(i)
                d.put_alt(ch, 0)
                d.ignored += a

    def most_popular_alternate(self):  # :107-120
        mx, best = F(0), " "
        if self.has_alternates():
            for i in range(6):
                c = self.alt_count_i(i)
                if c > mx or i == 0:
                    mx, best = c, KEYS[i]
        return best

    def counts_text(self, b, is_end):  # getCounts :122-149
        i = None if self.ref_char == b else KEYS.index(b)
        g = (lambda d: d.refc()) if i is None else (lambda d: d.alt(i))
        return format_number(g(self.c[(True, is_end)])) + "," + format_number(g(self.c[(False, is_end)]))


class Filter:  # MutationDetectionParameters
    def __init__(self, **kw):
        self.minSNPTotalDepth = self.minSNPDepthFraction = 0.0
        self.minIndelTotalStartDepth = self.minIndelStartDepthFraction = 0.0
        self.minIndelContinuationTotalDepth = self.minIndelContinuationDepthFraction = 0.0
        for k, v in kw.items():
            setattr(self, k, v)

    @staticmethod
    def default():
        return Filter(minSNPTotalDepth=5, minSNPDepthFraction=0.9, minIndelTotalStartDepth=1, minIndelStartDepthFraction=0.8,
                      minIndelContinuationTotalDepth=1, minIndelContinuationDepthFraction=0.7)

    def supports_snp(self, depth, total):
        if total < F(self.minSNPTotalDepth) or depth <= 0:
            return False
        return not (F(depth / total) < F(self.minSNPDepthFraction))

    def _indel(self, p, min_total, min_frac):
        mid = p.middle_count()
        if mid < F(min_total):
            return False
        if p.ref_char == "-":
            mi, ei = F(mid - p.middle_ref()), F(p.end_count() - p.end_ref())
        else:
            mi, ei = p.middle_alt_i(5), p.end_alt_i(5)
        if mi <= 0 and ei <= 0:
            return False
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = F(mi) / F(mid)
        return not (frac < F(min_frac))

    def supports_indel_start(self, p):
        return self._indel(p, self.minIndelTotalStartDepth, self.minIndelStartDepthFraction)

    def supports_indel_continuation(self, p):
        return self._indel(p, self.minIndelContinuationTotalDepth, self.minIndelContinuationDepthFraction)


class Store:
    """Map<Sequence, Alignments>: per contig the four DirectionalAlignments [region: 0 middle, 1 end][dir: 0 forward, 1 reverse]."""

    def __init__(self, contigs, end_fraction):
        """contigs: list of (name, uint8 codes) in database order (forward strands)."""
        self.contigs = contigs
        self.end_fraction = float(end_fraction)
        self.d = [[[Directional(c) for _ in range(2)] for _ in range(2)] for _, c in contigs]
        self.touched = [False] * len(contigs)  # MatchDatabase only holds an Alignments for contigs something aligned to

    # ---- Alignments.getPosition / getInsertion (AlignmentsSection :41-57, RegionAlignments :29-37, DirectionalAlignments :98-160) ----
    def position(self, contig, i):
        p = Position(LETTERS[int(self.contigs[contig][1][i])])
        for region, end in ((1, True), (0, False)):
            for dirn, forward in ((1, False), (0, True)):
                self._update_count(self.d[contig][region][dirn], p, i, forward, end)
        return p

    @staticmethod
    def _update_count(D, p, i, forward, end):
        va = D.alt.get(i)
        if va is not None:
            for v in va[0]:
                p.put_scaled(v.allele, v.count, forward, end)
                p.put_sample(v.ex, abs(v.ex_index), v.ex_index < 0)
        p.put_scaled(LETTERS[int(D.ref[i])], int(D.ref_counts[i]), forward, end)

    def insertion(self, contig, i, k):
        p = Position("-")
        for region, end in ((1, True), (0, False)):
            for dirn, forward in ((0, True), (1, False)):
                D = self.d[contig][region][dirn]
                scaled_ins = 0
                va = D.alt.get(i)
                if va is not None and len(va[1]) > k:
                    for v in va[1][k]:
                        p.put_scaled(v.allele, v.count, forward, end)
                        scaled_ins += v.count
                        p.put_sample(v.ex, v.ex_index, False)
                base = Position(LETTERS[int(D.ref[i])])
                self._update_count(D, base, i, forward, end)
                non = int(F(base.count() * F(100))) - scaled_ins
                if non > 0:
                    p.put_scaled("-", non, forward, end)
        return p

    # ---- FilteredAlignments ----
    def f_position(self, contig, i, flt):
        p = self.position(contig, i)
        total = p.count()
        for alt in p.nonzero_alternates():
            if alt != "-" and not flt.supports_snp(p.alt_count(alt), total):
                p.ignore_alternate(alt)
        if p.has_alternates() and not self._could_be_deletion(contig, i, flt):
            p.ignore_alternate("-")
        return p

    def _could_be_deletion(self, contig, i, flt):
        while i >= 0:
            p = self.position(contig, i)
            if flt.supports_indel_start(p):
                return True
            if not flt.supports_indel_continuation(p):
                return False
            i -= 1
        return True

    def f_insertion(self, contig, i, k, flt):
        p = self.insertion(contig, i, k)
        if not p.has_alternates():
            return p
        keep = flt.supports_indel_start(p) if k == 0 else flt.supports_indel_continuation(p)
        if not keep:
            for alt in p.nonzero_alternates():
                p.ignore_alternate(alt)
        return p

    # ---- the device's tables -> the same structure ----
    @staticmethod
    def from_device(contigs, end_fraction, planes, table, seq_lookup):
        """planes: per contig int32 [2][2][len] reference-base counts; table: dict of arrays (key fields decoded: contig, pos, region, dir,
        ins (-1: at the position, k: insertion column k), allele (index into KEYS), count, ex_gid, ex_rev, ex_index); seq_lookup(gid, rev) -> Seq."""
        S = Store(contigs, end_fraction)
        for c, pl in enumerate(planes):
            for region in range(2):
                for dirn in range(2):
                    S.d[c][region][dirn].ref_counts = pl[region, dirn].astype(np.int64)
            S.touched[c] = True
        for j in range(len(table["contig"])):
            D = S.d[int(table["contig"][j])][int(table["region"][j])][int(table["dir"][j])]
            va = D._vi(int(table["pos"][j]))
            k = int(table["ins"][j])
            if k < 0:
                lst = va[0]
            else:
                while len(va[1]) <= k:
                    va[1].append([])
                lst = va[1][k]
            v = D._get_or_create(lst, KEYS[int(table["allele"][j])])
            v.count += int(table["count"][j])
            v.ex = seq_lookup(int(table["ex_gid"][j]), int(table["ex_rev"][j]))
            v.ex_index = int(table["ex_index"][j])
        return S

    def variant_table(self):
        """Canonical listing of every variant: (contig, pos, region, dir, ins, allele index, count, example name, example id, example index)."""
        out = []
        for c in range(len(self.contigs)):
            for region in range(2):
                for dirn in range(2):
                    D = self.d[c][region][dirn]
                    for pos, va in D.alt.items():
                        for v in va[0]:
                            out.append((c, pos, region, dirn, -1, KEYS.index(v.allele), v.count, v.ex.name, v.ex.id, v.ex_index))
                        for k, col in enumerate(va[1]):
                            for v in col:
                                out.append((c, pos, region, dirn, k, KEYS.index(v.allele), v.count, v.ex.name, v.ex.id, v.ex_index))
        out.sort()
        return out


def accumulate(store, results, reads, names, first_seq_id=0):
    """MatchDatabase.addAlignments over one batch.  reads: per query a list of uint8 code arrays (mates as read); names: one name per
    SEQUENCE of the batch, in batch order; ids are first_seq_id + sequence index (Sequence ids follow the order the reads were parsed in)."""
    r = results
    sid0 = 0
    for q in range(len(reads)):
        n_mates = len(reads[q])
        seqs = []
        for m in range(n_mates):
            fwd = Seq(names[sid0 + m], first_seq_id + sid0 + m, reads[q][m])
            rev = Seq(names[sid0 + m] + "-rev", first_seq_id + sid0 + m, COMP[reads[q][m][::-1]])
            seqs.append((fwd, rev))
        sid0 += n_mates
        if r["q_status"][q] != 0:
            continue
        comps = range(r["q_comp_off"][q], r["q_comp_off"][q + 1])
        n_comp = len(comps)
        for ci, c in enumerate(comps):
            k0, k1 = int(r["comp_choice_off"][c]), int(r["comp_choice_off"][c + 1])
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

                def wgt(rb):  # WeightedAlignment.getWeight
                    num = n_sa
                    if n_sa >= 2 and (rb < min_ov or rb >= max_ov):
                        num = 1
                    pos_w = F(1.0) / F(num) if num != 0 else F(0)
                    return F(weight * pos_w)

                for si, s in enumerate(sas):
                    mate = ci if n_comp == 2 else si
                    rev = int(r["sa_reversed"][s])
                    A = seqs[mate][rev]
                    contig = int(r["sa_contig"][s])
                    store.touched[contig] = True
                    bl = spans[si]
                    first_start_a, last_end_a = int(bl[0, 0]), int(bl[-1, 0] + bl[-1, 2])
                    limit = len(A) * store.end_fraction

                    def near_end(qi):
                        return min(qi - first_start_a, last_end_a - qi - 1) < limit

                    for a0, b0, al, blen in bl.tolist():
                        if al == blen:
                            for i in range(al):
                                D = store.d[contig][1 if near_end(a0 + i) else 0][rev]
                                D.add(b0 + i, int(A.codes[a0 + i]), wgt(b0 + i), A, a0 + i)
                        elif al > blen:
                            D = store.d[contig][1 if near_end(a0) else 0][rev]
                            D.insert(b0 - 1, A.text(a0, al), wgt(b0 - 1), A, a0)
                        else:
                            for i in range(blen):
                                D = store.d[contig][1 if near_end(a0 + i) else 0][rev]
                                D.add(b0 + i, 0, wgt(b0 + i), A, a0)


# ---------------------------------------------------------------- formatters (bodies)
def _jobs(store):
    """VcfWriter/MutationsWriter.splitJobs: contigs that have an Alignments, sorted by name; job boundaries do not change the VCF
    body and only matter to the mutations body where a deletion run crosses one (jobs of 8192 positions, numParallelJobs = 1)."""
    order = sorted((name, c) for c, (name, _) in enumerate(store.contigs) if store.touched[c])
    for name, c in order:
        n = len(store.contigs[c][1])
        start = 0
        while start < n:
            end = min(n, start + 8192)
            yield name, c, start, end
            start = end


def mutations_body(store, flt=None):  # MutationsFormatterWorker.format :25-72
    flt = flt or Filter()
    out = []

    def write(name, row, ref, mut, depth, total):
        out.append("%s\t%d\t%s\t%s\t%s\t%s\n" % (name, row, ref, mut, format_number(depth), format_number(total)))

    def write_deletions(name, row, dels):
        depth = total = F(-1)
        for p in dels:
            t, d = p.middle_count(), p.middle_alt_i(5)
            if depth < 0 or depth > d:
                depth = d
            if total < 0 or total > t:
                total = t
        write(name, row, "".join(p.ref_char for p in dels), "-" * len(dels), depth, total)

    for name, c, start, end in _jobs(store):
        dels = []
        display = 1
        for i in range(start, end):
            display = i + 1
            p = store.f_position(c, i, flt)
            if p.count() > 0:
                if p.has_alternates() and p.ref_char != "N":
                    for alt in p.nonzero_alternates():
                        if alt != "-":
                            write(name, display, p.ref_char, alt, p.alt_count(alt), p.count())
                cand = []
                k = 0
                while True:
                    ins = store.f_insertion(c, i, k, flt)
                    if not ins.has_alternates():
                        break
                    cand.append(ins)
                    k += 1
                if cand:
                    total = p.middle_count()
                    depth = total
                    for ins in cand:
                        depth = F(ins.middle_count() - ins.middle_ref())
                    write(name, display, "-" * len(cand), "".join(x.most_popular_alternate() for x in cand), depth, total)
            if p.alt_count("-") > 0:
                dels.append(p)
            elif dels:
                write_deletions(name, display - len(dels), dels)
                dels = []
        if dels:
            write_deletions(name, display - (len(dels) - 1), dels)
    return "".join(out)


def vcf_body(store, flt=None, include_non_mutations=True, show_support=True):  # VcfFormatterWorker.format :27-110
    flt = flt or Filter()
    out = []

    def field(s):
        return s if s else "."

    def write(name, row, p):
        ref = p.ref_char
        alts = p.nonzero_alternates()
        cols = [name, str(row), ref, field(",".join(alts)), format_number(p.count())]
        for is_end in (False, True):
            cols.append(";".join([p.counts_text(ref, is_end)] + [p.counts_text(a, is_end) for a in alts]))
        if show_support:
            parts = []
            for a in alts:
                s = p.sample[KEYS.index(a)] if p.sample is not None else None
                if s is None:
                    parts.append("")
                    continue
                seq, idx = s
                if idx < 0:
                    idx = -idx
                    parts.append(seq.text(0, idx) + "[-]" + seq.text(idx, len(seq) - idx))
                else:
                    parts.append(seq.text(0, idx) + "[" + LETTERS[int(seq.codes[idx])] + "]" + seq.text(idx + 1, len(seq) - 1 - idx))
            cols.append(field(",".join(parts)))
        out.append("\t".join(cols) + "\n")

    for name, c, start, end in _jobs(store):
        for i in range(start, end):
            p = store.f_position(c, i, flt)
            if p.count() > 0 and (include_non_mutations or p.has_alternates()):
                write(name, i + 1, p)
            k = 0
            while True:
                ins = store.f_insertion(c, i, k, flt)
                if not ins.has_alternates():
                    break
                write(name, -(i + 1), ins)
                k += 1
    return "".join(out)
