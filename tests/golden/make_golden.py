#!/usr/bin/env python
"""Generates tests/golden/rows_v1.tsv: the 15-column bwa_result rows (plus MAPQ and NM, which the path
computes but the SQL row does not export) for a small deterministic workload, produced by the CPU oracle.
The reference itself cannot be run here (no PostgreSQL / libbwa), so these vectors freeze the oracle's
behaviour -- they pin regressions, not a libbwa binary ("parity unpinned").
    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def workload():
    import numpy as np
    from bioseqdb_b200 import synth
    rows = synth.reference_rows([30_001, 20_002, 1_503], seed=1234)
    rows = synth.plant_repeats(rows, n_families=3, copies=4, unit=(200, 400), divergence=0.02, seed=1235)
    rows[1][5000:5012] = np.frombuffer(b"NNNNNNRRYYKM", dtype=np.uint8)   # holes in the reference
    seqs, offs, _ = synth.simulate_reads(rows, 160, 150, sub=0.02, ins=0.003, dele=0.003, seed=1236, n_frac=0.003)
    ids = synth.lrand48_ids_fast(160)
    return rows, seqs, offs, ids


def render(rows_by_read):
    out = []
    for qi, ms in enumerate(rows_by_read):
        for m in ms:
            out.append("\t".join(str(x) for x in (m["ref_id"], m["ref_subseq"], m["ref_match_begin"], m["ref_match_end"], m["ref_match_len"], qi + 1,
                                                  m["query_subseq"], m["query_match_begin"], m["query_match_end"], m["query_match_len"],
                                                  "t" if m["is_primary"] else "f", "t" if m["is_secondary"] else "f", "t" if m["is_reverse"] else "f",
                                                  m["cigar"], m["score"], m["mapq"], m["nm"])))
    return "\n".join(out) + "\n"


def oracle_rows():
    import oracle_lib as O
    rows, seqs, offs, ids = workload()
    ix = O.OracleIndex(O.sql_default_opts(len(rows)))
    for i, r in enumerate(rows):
        ix.add_ref_text(100 + i, r.tobytes())
    ix.build()
    res = ix.align_batch(seqs, offs, ids, 1)
    out = []
    for i in range(len(offs) - 1):
        read = seqs[int(offs[i]):int(offs[i + 1])].tobytes()
        txt = ix.rows_text(read, int(ids[i]))
        rws = res["rows"][int(res["row_off"][i]):int(res["row_off"][i + 1])]
        ms = []
        for line, rw in zip(txt.strip("\n").split("\n") if txt.strip() else [], rws):
            f = line.split("\t")
            ms.append(dict(ref_id=int(f[0]), ref_subseq=f[1], ref_match_begin=int(f[2]), ref_match_end=int(f[3]), ref_match_len=int(f[4]),
                           query_subseq=f[5], query_match_begin=int(f[6]), query_match_end=int(f[7]), query_match_len=int(f[8]),
                           is_primary=f[9] == "t", is_secondary=f[10] == "t", is_reverse=f[11] == "t", cigar=f[12], score=int(f[13]),
                           mapq=int(rw["mapq"]), nm=int(rw["NM"])))
        out.append(ms)
    return out


if __name__ == "__main__":
    txt = render(oracle_rows())
    with open(os.path.join(HERE, "rows_v1.tsv"), "w") as f:
        f.write(txt)
    print("wrote %d rows" % txt.count("\n"))
