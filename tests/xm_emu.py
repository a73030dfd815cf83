"""ctypes binding of the host emulation harness (tests/emu/libxmemu.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

RESULT_ARRAYS = [("q_comp_off", np.int64), ("comp_choice_off", np.int64), ("choice_sa_off", np.int64), ("sa_block_off", np.int64),
                 ("choice_f64", np.float64), ("sa_f64", np.float64), ("choice_inner", np.int32), ("sa_contig", np.int32),
                 ("blocks", np.int32), ("q_status", np.int32), ("sa_reversed", np.uint8), ("stats", np.int64)]


class CParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mutation", "ins_start", "ins_ext", "del_start", "del_ext", "max_error_rate",
                                           "unaligned", "ambiguity", "max_penalty_span")] + [("max_num_matches", C.c_int32), ("enable_gapmers", C.c_int32)]


def make_params(d):
    p = CParams()
    for k in ("mutation", "ins_start", "ins_ext", "del_start", "del_ext", "max_error_rate", "unaligned", "ambiguity", "max_penalty_span"):
        setattr(p, k, float(d[k]))
    p.max_num_matches = int(d.get("max_num_matches", 2147483647))
    p.enable_gapmers = int(d.get("enable_gapmers", 1))
    return p


def lib():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
        L = C.CDLL(os.path.join(ROOT, "tests", "emu", "libxmemu.so"))
        L.xe_create.restype = C.c_void_p
        L.xe_last_error.restype = C.c_char_p
        L.xe_align_batch.restype = C.c_void_p
        L.xe_results_array.restype = C.c_int64
        _LIB = L
    return _LIB


def read_results(getter, r):
    out = {}
    for i, (name, dt) in enumerate(RESULT_ARRAYS):
        ptr = C.c_void_p()
        n = getter(C.c_void_p(r), i, C.byref(ptr))
        if n > 0:
            out[name] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()
        else:
            out[name] = np.zeros(0, dtype=dt)
    return out


class Emu:
    def __init__(self, params):
        self.L = lib()
        p = make_params(params)
        self.h = self.L.xe_create(C.byref(p))
        self._keep = []

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.xe_last_error(C.c_void_p(self.h)).decode())

    def set_reference(self, packed_contigs, lengths):
        n = len(packed_contigs)
        arr = (C.c_void_p * n)(*[p.ctypes.data for p in packed_contigs])
        lens = np.asarray(lengths, dtype=np.int32)
        self._keep = [packed_contigs, lens]
        self._ok(self.L.xe_set_reference(C.c_void_p(self.h), n, arr, lens.ctypes.data_as(C.c_void_p)))

    def set_position_bias(self, bias):
        """Global position of the first contig (call before set_reference): moves every index position past 2^32."""
        self._ok(self.L.xe_set_position_bias(C.c_void_p(self.h), C.c_longlong(int(bias))))

    def set_index_length(self, t, wide=False):
        off = np.ascontiguousarray(t["offsets"], dtype=np.int64)
        over = np.ascontiguousarray(t["overfull"], dtype=np.uint8)
        pos = np.ascontiguousarray(t["positions"], dtype=np.uint64 if wide else np.uint32)
        if len(pos) == 0:
            pos = np.zeros(1, dtype=pos.dtype)
        f = self.L.xe_set_index_length_wide if wide else self.L.xe_set_index_length
        self._ok(f(C.c_void_p(self.h), t["used"], t["capacity"], t["max_count"], off.ctypes.data_as(C.c_void_p),
                   over.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p)))

    def finish_index(self, min_interesting, max_built):
        self._ok(self.L.xe_finish_index(C.c_void_p(self.h), min_interesting, max_built))

    def build_index(self, max_used, threads=1):
        self._ok(self.L.xe_build_index(C.c_void_p(self.h), max_used, threads))

    def index_info(self):
        a, b = C.c_int(), C.c_int()
        self.L.xe_index_info(C.c_void_p(self.h), C.byref(a), C.byref(b))
        return a.value, b.value

    def get_index_length(self, n, wide=False):
        cap, mx, npos = C.c_int(), C.c_int(), C.c_int64()
        self._ok(self.L.xe_get_index_length(C.c_void_p(self.h), n, C.byref(cap), C.byref(mx), C.byref(npos), None, None, None))
        off = np.zeros(cap.value + 1, dtype=np.int64)
        over = np.zeros(cap.value, dtype=np.uint8)
        pos = np.zeros(max(npos.value, 1), dtype=np.uint32)
        self._ok(self.L.xe_get_index_length(C.c_void_p(self.h), n, C.byref(cap), C.byref(mx), C.byref(npos), off.ctypes.data_as(C.c_void_p),
                                            over.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p)))
        t = dict(used=n, capacity=cap.value, max_count=mx.value, offsets=off, overfull=over, positions=pos[:npos.value])
        if wide:
            pos64 = np.zeros(max(npos.value, 1), dtype=np.uint64)
            self._ok(self.L.xe_get_index_positions_wide(C.c_void_p(self.h), n, pos64.ctypes.data_as(C.c_void_p)))
            t["positions"] = pos64[:npos.value]
        return t

    def set_duplications(self, window, granularity, contig, starts):
        s = np.ascontiguousarray(starts, dtype=np.int32)
        if len(s) == 0:
            s = np.zeros(1, dtype=np.int32)
        self._ok(self.L.xe_set_duplications(C.c_void_p(self.h), window, C.c_double(granularity), contig, len(starts), s.ctypes.data_as(C.c_void_p)))

    def build_duplications(self, min_len=-1, max_len=-1, min_copies=2, window=1000, via_merge=False):
        """via_merge: the found blocks go through HostModel::merge_duplications, the routine the device scan feeds."""
        f = self.L.xe_build_duplications_via_merge if via_merge else self.L.xe_build_duplications
        self._ok(f(C.c_void_p(self.h), min_len, max_len, min_copies, window))

    def get_duplications(self, contig):
        n = C.c_int()
        self.L.xe_get_duplications(C.c_void_p(self.h), contig, C.byref(n), None)
        out = np.zeros(max(n.value, 1), dtype=np.int32)
        self.L.xe_get_duplications(C.c_void_p(self.h), contig, C.byref(n), out.ctypes.data_as(C.c_void_p))
        return out[:n.value]

    def align_batch(self, batch, threads=1, max_tier=2):
        nq = len(batch["n_seqs"])
        r = self.L.xe_align_batch(C.c_void_p(self.h), nq, batch["packed"].ctypes.data_as(C.c_void_p), batch["seq_word_off"].ctypes.data_as(C.c_void_p),
                                  batch["seq_len"].ctypes.data_as(C.c_void_p), batch["n_seqs"].ctypes.data_as(C.c_void_p),
                                  batch["expected_inner"].ctypes.data_as(C.c_void_p), batch["per_penalty"].ctypes.data_as(C.c_void_p), threads, max_tier)
        out = read_results(self.L.xe_results_array, r)
        self.L.xe_results_free(C.c_void_p(r))
        return out

    def close(self):
        if self.h:
            self.L.xe_destroy(C.c_void_p(self.h))
            self.h = None
