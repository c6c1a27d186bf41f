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


def _flat_cigars(rows, cigar):
    nc = rows["n_cigar"].astype(np.int64)
    tot = int(nc.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.uint32), nc
    starts = np.repeat(rows["cigar_off"].astype(np.int64) - np.concatenate(([0], np.cumsum(nc)[:-1])), nc)
    return cigar[starts + np.arange(tot, dtype=np.int64)], nc


def parity_report(gres, ores, n_reads=None):
    """Vectorised read-level comparison of GPU rows with oracle rows over the first n_reads reads: every PARITY_FIELDS column and the
    CIGAR words.  A read mismatches when its row count, any field of any of its rows, or any CIGAR word differs.
    Returns {"reads_checked", "mismatching_reads", "rows_checked", "fields", "first_mismatching_reads"}."""
    g_off = gres.row_off.astype(np.int64)
    o_off = ores["row_off"].astype(np.int64)
    n = int(n_reads if n_reads is not None else min(len(g_off), len(o_off)) - 1)
    g_cnt, o_cnt = np.diff(g_off[:n + 1]), np.diff(o_off[:n + 1])
    bad = g_cnt != o_cnt
    same = ~bad
    g_rows = gres.rows[int(g_off[0]):int(g_off[n])][np.repeat(same, g_cnt)]
    o_rows = ores["rows"][int(o_off[0]):int(o_off[n])][np.repeat(same, o_cnt)]
    read_of_row = np.repeat(np.arange(n)[same], g_cnt[same])
    row_bad = np.zeros(len(g_rows), dtype=bool)
    for f in PARITY_FIELDS:
        row_bad |= g_rows[f] != o_rows[f]
    gc, gn = _flat_cigars(g_rows, gres.cigar)
    oc, on = _flat_cigars(o_rows, ores["cigar"])
    if len(gc) == len(oc) and np.array_equal(gn, on):
        if len(gc):
            word_bad = gc != oc
            if word_bad.any():
                row_bad[np.unique(np.repeat(np.arange(len(g_rows)), gn)[word_bad])] = True
    else:   # n_cigar differs somewhere: those rows are already flagged by the n_cigar field; compare the rest row by row
        eq = gn == on
        for i in np.flatnonzero(eq & ~row_bad):
            a = gres.cigar[int(g_rows["cigar_off"][i]):int(g_rows["cigar_off"][i]) + int(gn[i])]
            b = ores["cigar"][int(o_rows["cigar_off"][i]):int(o_rows["cigar_off"][i]) + int(on[i])]
            if not np.array_equal(a, b):
                row_bad[i] = True
    bad[np.unique(read_of_row[row_bad])] = True
    return {"reads_checked": n, "mismatching_reads": int(bad.sum()), "rows_checked": int(len(g_rows)), "fields": len(PARITY_FIELDS),
            "cigars_compared": True, "first_mismatching_reads": np.flatnonzero(bad)[:8].tolist()}
