"""ctypes binding of the oracle (oracle/libxmoracle.so). TEST INFRASTRUCTURE ONLY — never imported by mapper_b200."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class CParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mutation", "ins_start", "ins_ext", "del_start", "del_ext", "max_error_rate",
                                           "unaligned", "ambiguity", "max_penalty_span")] + [("max_num_matches", C.c_int32), ("reserved", C.c_int32)]


def make_params(d):
    p = CParams()
    for k in ("mutation", "ins_start", "ins_ext", "del_start", "del_ext", "max_error_rate", "unaligned", "ambiguity", "max_penalty_span"):
        setattr(p, k, float(d[k]))
    p.max_num_matches = int(d.get("max_num_matches", 2147483647))
    return p


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libxmoracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        L = C.CDLL(path)
        L.xo_create.restype = C.c_void_p
        L.xo_last_error.restype = C.c_char_p
        L.xo_contig_name.restype = C.c_char_p
        L.xo_dup_granularity.restype = C.c_double
        L.xo_dup_count.restype = C.c_int64
        L.xo_align_batch.restype = C.c_void_p
        L.xo_results_array.restype = C.c_int64
        for f in ("xo_align_json", "xo_test_path_aligner", "xo_test_hashblock_aligner", "xo_test_counting_path", "xo_test_paths_counter"):
            getattr(L, f).restype = C.c_void_p
        _LIB = L
    return _LIB


def _take(ptr, ctx=None):
    L = lib()
    if not ptr:
        raise RuntimeError("oracle error: " + (L.xo_last_error(C.c_void_p(ctx)).decode() if ctx else "null"))
    s = C.string_at(ptr).decode()
    L.xo_free(C.c_void_p(ptr))
    return s


class Oracle:
    """One reference + index + duplication detector (mirrors Api.newDatabase / Mapper.run set-up)."""

    def __init__(self, contigs, sort_by_length=False, min_interesting=-1, max_short=-1, gapmers=True, threads=1,
                 dup=None):
        L = lib()
        self.L = L
        self.h = L.xo_create()
        for name, text in contigs:
            assert L.xo_add_contig(C.c_void_p(self.h), name.encode(), text.encode()) == 0
        self._ok(L.xo_finalize_reference(C.c_void_p(self.h), int(sort_by_length)))
        self._ok(L.xo_create_index(C.c_void_p(self.h), min_interesting, max_short, int(gapmers), threads))
        d = dup or {}
        self._ok(L.xo_create_dup_detector(C.c_void_p(self.h), d.get("min_len", -1), d.get("max_len", -1), d.get("min_copies", 2), d.get("window", 1)))

    def _ok(self, rc):
        if rc < 0:
            raise RuntimeError("oracle error: " + self.L.xo_last_error(C.c_void_p(self.h)).decode())
        return rc

    def close(self):
        if self.h:
            self.L.xo_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference / index export (plays the Java host's role when feeding the product C-ABI) ----
    def num_contigs(self):
        return self.L.xo_num_contigs(C.c_void_p(self.h))

    def contig(self, i):
        n = self.L.xo_contig_length(C.c_void_p(self.h), i)
        buf = np.empty(n, dtype=np.uint8)
        self.L.xo_contig_codes(C.c_void_p(self.h), i, buf.ctypes.data_as(C.c_void_p))
        return self.L.xo_contig_name(C.c_void_p(self.h), i).decode(), buf

    def build_through(self, n):
        return self._ok(self.L.xo_build_index_through(C.c_void_p(self.h), int(n)))

    def min_interesting(self):
        return self.L.xo_index_min_interesting(C.c_void_p(self.h))

    def max_built(self):
        return self.L.xo_index_max_built(C.c_void_p(self.h))

    def table(self, n):
        cap, mx, npos = C.c_int(), C.c_int(), C.c_int64()
        self._ok(self.L.xo_index_table_info(C.c_void_p(self.h), n, C.byref(cap), C.byref(mx), C.byref(npos)))
        offsets = np.zeros(cap.value + 1, dtype=np.int64)
        positions = np.zeros(max(npos.value, 1), dtype=np.uint32)
        overfull = np.zeros(cap.value, dtype=np.uint8)
        self.L.xo_index_table_copy(C.c_void_p(self.h), n, offsets.ctypes.data_as(C.c_void_p), positions.ctypes.data_as(C.c_void_p), overfull.ctypes.data_as(C.c_void_p))
        return dict(used=n, capacity=cap.value, max_count=mx.value, offsets=offsets, positions=positions[:npos.value], overfull=overfull)

    def detect_duplications(self):
        self._ok(self.L.xo_dup_detect(C.c_void_p(self.h)))

    def dup_granularity(self):
        return self.L.xo_dup_granularity(C.c_void_p(self.h))

    def dup_starts(self, contig):
        n = self.L.xo_dup_count(C.c_void_p(self.h), contig)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.L.xo_dup_copy(C.c_void_p(self.h), contig, out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def infer_ancestors(self, dissimilarity_threshold, verify=False):
        """AncestryDetector.unionRecentAncestors over this reference (the Oracle must have been created with dup=dict(min_copies=3, window=1),
        M/Mapper.java:675-681).  Returns [(name + "-anc", codes)] in database order."""
        n = self._ok(self.L.xo_infer_ancestors(C.c_void_p(self.h), C.c_double(dissimilarity_threshold), int(verify)))
        out = []
        for i in range(n):
            name, codes = self.contig(i)
            buf = np.empty(len(codes), dtype=np.uint8)
            self.L.xo_ancestor_codes(C.c_void_p(self.h), i, buf.ctypes.data_as(C.c_void_p))
            out.append((name + "-anc", buf))
        return out

    # ---- alignment ----
    def align(self, params, seqs, expected_inner=0.0, per_penalty=1.0):
        p = make_params(params)
        s2 = seqs[1].encode() if len(seqs) > 1 else None
        ptr = self.L.xo_align_json(C.c_void_p(self.h), C.byref(p), seqs[0].encode(), s2, C.c_double(expected_inner), C.c_double(per_penalty))
        return json.loads(_take(ptr, self.h))

    def align_batch(self, params, batch, threads=1):
        """batch: dict(packed uint16, seq_word_off int64, seq_len int32, n_seqs uint8, expected_inner f64, per_penalty f64)."""
        p = make_params(params)
        nq = len(batch["n_seqs"])
        r = self.L.xo_align_batch(C.c_void_p(self.h), C.byref(p), nq,
                                  batch["packed"].ctypes.data_as(C.c_void_p), batch["seq_word_off"].ctypes.data_as(C.c_void_p),
                                  batch["seq_len"].ctypes.data_as(C.c_void_p), batch["n_seqs"].ctypes.data_as(C.c_void_p),
                                  batch["expected_inner"].ctypes.data_as(C.c_void_p), batch["per_penalty"].ctypes.data_as(C.c_void_p), threads)
        if not r:
            raise RuntimeError("oracle error: " + self.L.xo_last_error(C.c_void_p(self.h)).decode())
        names = [("q_comp_off", np.int64), ("comp_choice_off", np.int64), ("choice_sa_off", np.int64), ("sa_block_off", np.int64),
                 ("choice_f64", np.float64), ("sa_f64", np.float64), ("choice_inner", np.int32), ("sa_contig", np.int32),
                 ("blocks", np.int32), ("q_status", np.int32), ("sa_reversed", np.uint8), ("stats", np.int64)]
        out = {}
        for i, (name, dt) in enumerate(names):
            ptr = C.c_void_p()
            n = self.L.xo_results_array(C.c_void_p(r), i, C.byref(ptr))
            if n > 0:
                arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()
            else:
                arr = np.zeros(0, dtype=dt)
            out[name] = arr
        self.L.xo_results_free(C.c_void_p(r))
        return out

    def counting_path(self, params, query, priority):
        p = make_params(params)
        return json.loads(_take(self.L.xo_test_counting_path(C.c_void_p(self.h), C.byref(p), query.encode(), priority), self.h))

    def paths_counter(self, params, seq1, seq2, expected_inner, max_inner):
        p = make_params(params)
        return json.loads(_take(self.L.xo_test_paths_counter(C.c_void_p(self.h), C.byref(p), seq1.encode(), seq2.encode(), expected_inner, max_inner), self.h))


def path_aligner(params, a, b, max_ins, max_del):
    p = make_params(params)
    s = _take(lib().xo_test_path_aligner(C.byref(p), a.encode(), b.encode(), C.c_double(max_ins), C.c_double(max_del)))
    if s.startswith("error:"):
        raise RuntimeError(s)
    return None if s == "null" else json.loads(s)


def hashblock_aligner(params, a, b, max_ins, max_del):
    p = make_params(params)
    s = _take(lib().xo_test_hashblock_aligner(C.byref(p), a.encode(), b.encode(), C.c_double(max_ins), C.c_double(max_del)))
    if s.startswith("error:"):
        raise RuntimeError(s)
    return None if s == "null" else json.loads(s)


def hash_symmetry(text):
    return lib().xo_test_hash_symmetry(text.encode())
