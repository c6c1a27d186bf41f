"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

ROW_DTYPE = np.dtype([
    ("rb", "<i8"), ("re", "<i8"), ("pos", "<i8"), ("hash", "<u8"),
    ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"), ("truesc", "<i4"), ("sub", "<i4"),
    ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"), ("seedcov", "<i4"), ("secondary", "<i4"),
    ("seedlen0", "<i4"), ("n_comp", "<i4"), ("frac_rep", "<f4"),
    ("is_rev", "<i4"), ("mapq", "<i4"), ("NM", "<i4"), ("flag", "<i4"),
    ("cigar_off", "<u4"), ("n_cigar", "<u4"), ("ref_id", "<i8"),
])
assert ROW_DTYPE.itemsize == 120

HOLE_DTYPE = np.dtype([("offset", "<i8"), ("len", "<i4"), ("amb", "S1"), ("_pad", "V3")])
assert HOLE_DTYPE.itemsize == 16


class Opts(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "min_seed_len", "max_occ", "a", "b", "pen_clip3", "pen_clip5", "zdrop", "w", "o_del", "e_del", "o_ins", "e_ins")]


def sql_default_opts(n_rows: int = 1) -> Opts:
    """What bwa_opts() really delivers: o_del 6, e_del 6, o_ins 1, e_ins 1."""
    return Opts(19, max(500, 2 * n_rows), 1, 4, 5, 5, 100, 100, 6, 6, 1, 1)


def canonical_opts(n_rows: int = 1) -> Opts:
    return Opts(19, max(500, 2 * n_rows), 1, 4, 5, 5, 100, 100, 6, 1, 6, 1)


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.POINTER(Opts)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_opts.argtypes = [C.c_void_p, C.POINTER(Opts)]
        L.orc_add_ref_text.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_uint64]
        L.orc_add_ref_packed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_build.argtypes = [C.c_void_p]
        L.orc_build.restype = C.c_double
        L.orc_adopt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_index_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_index_bwt.restype = C.c_void_p
        L.orc_index_bwt.argtypes = [C.c_void_p]
        L.orc_index_sa.restype = C.c_void_p
        L.orc_index_sa.argtypes = [C.c_void_p]
        L.orc_index_pac.restype = C.c_void_p
        L.orc_index_pac.argtypes = [C.c_void_p]
        L.orc_index_anns.argtypes = [C.c_void_p] * 4
        L.orc_index_bwt_plain.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_suffix_array.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_bwt_sa.restype = C.c_uint64
        L.orc_bwt_sa.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_bwt_occ4.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_nuclseq_from_text.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_nuclseq_to_text.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_lrand48.argtypes = [C.c_int, C.c_void_p]
        L.orc_minstd.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_hash64.restype = C.c_uint64
        L.orc_hash64.argtypes = [C.c_uint64]
        L.orc_introsort_u64.argtypes = [C.c_uint64, C.c_void_p]
        L.orc_ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Opts), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_ksw_global2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Opts), C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_ksw_local.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Opts)]
        L.orc_stage_dump.restype = C.c_int64
        L.orc_stage_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                     C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64]
        L.orc_align_batch.restype = C.c_void_p
        L.orc_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_result_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_result_seconds.restype = C.c_double
        L.orc_result_seconds.argtypes = [C.c_void_p]
        for f in ("orc_result_row_off", "orc_result_rows", "orc_result_cigar"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_align_rows_text.restype = C.c_int64
        L.orc_align_rows_text.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_int64]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


COUNTER_NAMES = ["n_extend", "n_lf", "n_sa", "ext_cells", "ext_calls", "ext_rows", "glb_cells", "glb_calls",
                 "sw_cells", "sw_calls", "dup_chain_pos"]


class OracleIndex:
    """Mirror of the reference's BwaIndex (bwa.h:32-48) on the oracle."""

    def __init__(self, opts: Opts | None = None):
        self.L = lib()
        self.opts = opts or sql_default_opts()
        self.h = self.L.orc_new(C.byref(self.opts))
        self.n_rows = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_free(self.h)
            self.h = None

    def set_opts(self, opts: Opts):
        self.opts = opts
        self.L.orc_set_opts(self.h, C.byref(opts))

    def add_ref_text(self, rid: int, text: bytes):
        rc = self.L.orc_add_ref_text(self.h, rid, text, len(text))
        if rc:
            raise ValueError("invalid nucleotide in nuclseq_in: '%s'" % chr(rc if rc > 0 else 0))
        self.n_rows += 1

    def build(self) -> float:
        return self.L.orc_build(self.h)

    def adopt(self, plain_bwt: np.ndarray, primary: int, sa: np.ndarray):
        plain_bwt = np.ascontiguousarray(plain_bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(sa, dtype=np.uint64)
        self.L.orc_adopt(self.h, _ptr(plain_bwt), primary, _ptr(sa))

    def info(self):
        v = np.zeros(12, dtype=np.uint64)
        self.L.orc_index_info(self.h, _ptr(v))
        return dict(l_pac=int(v[0]), seq_len=int(v[1]), primary=int(v[2]), L2=[int(x) for x in v[3:8]],
                    bwt_size=int(v[8]), n_sa=int(v[9]), n_anns=int(v[10]), n_holes=int(v[11]))

    def bwt_plain(self) -> np.ndarray:
        n = (self.info()["seq_len"] + 15) // 16
        out = np.zeros(n, dtype=np.uint32)
        self.L.orc_index_bwt_plain(self.h, _ptr(out))
        return out

    def bwt_interleaved(self) -> np.ndarray:
        n = self.info()["bwt_size"]
        return np.ctypeslib.as_array(C.cast(self.L.orc_index_bwt(self.h), C.POINTER(C.c_uint32)), shape=(n,)).copy()

    def sa(self) -> np.ndarray:
        n = self.info()["n_sa"]
        return np.ctypeslib.as_array(C.cast(self.L.orc_index_sa(self.h), C.POINTER(C.c_uint64)), shape=(n,)).copy()

    def pac(self) -> np.ndarray:
        n = self.info()["l_pac"] // 4
        return np.ctypeslib.as_array(C.cast(self.L.orc_index_pac(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def bwt_sa(self, k: int) -> int:
        return self.L.orc_bwt_sa(self.h, k)

    def align_batch(self, seqs: np.ndarray, offs: np.ndarray, ids: np.ndarray, n_threads: int = 1):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = len(offs) - 1
        r = self.L.orc_align_batch(self.h, _ptr(seqs), _ptr(offs), _ptr(ids), n, n_threads)
        try:
            info = np.zeros(14, dtype=np.uint64)
            self.L.orc_result_info(r, _ptr(info))
            n_rows, n_cig = int(info[0]), int(info[1])
            row_off = np.ctypeslib.as_array(C.cast(self.L.orc_result_row_off(r), C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
            rows = np.zeros(n_rows, dtype=ROW_DTYPE)
            if n_rows:
                C.memmove(_ptr(rows), self.L.orc_result_rows(r), n_rows * ROW_DTYPE.itemsize)
            cig = np.zeros(n_cig, dtype=np.uint32)
            if n_cig:
                C.memmove(_ptr(cig), self.L.orc_result_cigar(r), n_cig * 4)
            ctr = {k: int(v) for k, v in zip(COUNTER_NAMES, info[2:13])}
            return dict(row_off=row_off, rows=rows, cigar=cig, counters=ctr, seconds=self.L.orc_result_seconds(r))
        finally:
            self.L.orc_result_free(r)

    def rows_text(self, seq: bytes, rid: int = 0) -> str:
        buf = C.create_string_buffer(1 << 20)
        n = self.L.orc_align_rows_text(self.h, seq, len(seq), rid, buf, len(buf))
        assert n >= 0
        return buf.value.decode()

    def stage_dump(self, seq: bytes):
        iv = np.zeros((4096, 4), dtype=np.uint64)
        sd = np.zeros((65536, 3), dtype=np.int64)
        ch = np.zeros(1 << 20, dtype=np.int64)
        ni, ns = C.c_int64(), C.c_int64()
        k = self.L.orc_stage_dump(self.h, seq, len(seq), _ptr(iv), len(iv), C.byref(ni), _ptr(sd), len(sd), C.byref(ns), _ptr(ch), len(ch))
        assert k >= 0 and ni.value <= len(iv) and ns.value <= len(sd)
        chains = []
        p = 0
        while p < k:
            pos, rid, n, w, kept = [int(x) for x in ch[p:p + 5]]
            p += 5
            seeds = ch[p:p + 3 * n].reshape(n, 3).copy()
            p += 3 * n
            chains.append(dict(pos=pos, rid=rid, w=w, kept=kept, seeds=seeds))
        return iv[:ni.value].copy(), sd[:ns.value].copy(), chains


def cigar_str(cig: np.ndarray) -> str:
    """htslib letters applied to bwa op codes (reference bwa.cpp:70-77; soft clip prints as N)."""
    return "".join("%d%s" % (int(c) >> 4, "MIDNSHP=XB"[int(c) & 0xF]) for c in cig)
