"""Shared helpers for the parity tests: build the same reference in the oracle and in the GPU library,
compare result rows field by field."""
import ctypes as C

import numpy as np

import oracle_lib as O
from bioseqdb_b200 import BwaIndex, BsqOpts, synth

PARITY_FIELDS = ["rid", "rb", "re", "qb", "qe", "is_rev", "score", "truesc", "secondary", "sub", "sub_n", "seedcov", "seedlen0", "w",
                 "n_comp", "frac_rep", "hash", "pos", "NM", "mapq", "flag", "n_cigar", "ref_id", "csub"]


def to_bsq(o: O.Opts) -> BsqOpts:
    return BsqOpts(*[getattr(o, f[0]) for f in O.Opts._fields_])


def build_pair(rows, opts: O.Opts, device=0, ids=None):
    orc = O.OracleIndex(opts)
    gpu = BwaIndex(device, to_bsq(opts))
    for i, r in enumerate(rows):
        rid = ids[i] if ids is not None else i + 1
        text = r.tobytes() if isinstance(r, np.ndarray) else bytes(r)
        orc.add_ref_text(rid, text)
        gpu.add_ref_sequence(rid, text)
    orc.build()
    gpu.build()
    return orc, gpu


def compare_results(gres, ores, max_report=5):
    """Returns a list of mismatch descriptions (empty = bit-exact)."""
    bad = []
    if not np.array_equal(gres.row_off, ores["row_off"]):
        d = np.flatnonzero(np.diff(gres.row_off.astype(np.int64)) != np.diff(ores["row_off"].astype(np.int64)))
        bad.append("row counts differ for %d reads, first %s" % (len(d), d[:max_report].tolist()))
        return bad
    for f in PARITY_FIELDS:
        a, b = gres.rows[f], ores["rows"][f]
        if not np.array_equal(a, b):
            d = np.flatnonzero(a != b)
            bad.append("field %s differs in %d rows, first rows %s gpu=%s oracle=%s" % (f, len(d), d[:max_report].tolist(), a[d[:max_report]].tolist(), b[d[:max_report]].tolist()))
    # cigars
    if not bad:
        go, oo = gres.rows["cigar_off"], ores["rows"]["cigar_off"]
        n = gres.rows["n_cigar"]
        for i in range(len(n)):
            ga = gres.cigar[int(go[i]):int(go[i]) + int(n[i])]
            oa = ores["cigar"][int(oo[i]):int(oo[i]) + int(n[i])]
            if not np.array_equal(ga, oa):
                bad.append("cigar differs at row %d: gpu=%s oracle=%s" % (i, O.cigar_str(ga), O.cigar_str(oa)))
                if len(bad) >= max_report:
                    break
    return bad


def read_arrays(reads):
    """list of bytes -> (seqs, offs)"""
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    seqs = np.frombuffer(b"".join(reads), dtype=np.uint8) if reads else np.zeros(0, dtype=np.uint8)
    return seqs, offs
