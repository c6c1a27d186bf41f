"""Committed golden rows (tests/golden/rows_v1.tsv, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the GPU path must produce the same 15-column rows through
the Python mirror of BwaIndex::align_sequence (GPU)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLDEN = open(os.path.join(HERE, "golden", "rows_v1.tsv")).read()


def test_oracle_reproduces_golden(oracle):
    assert make_golden.render(make_golden.oracle_rows()) == GOLDEN


@pytest.mark.gpu
def test_gpu_reproduces_golden(gpu_lib):
    import oracle_lib as O
    from bioseqdb_b200 import BwaIndex
    from helpers import to_bsq
    rows, seqs, offs, ids = make_golden.workload()
    ix = BwaIndex(0, to_bsq(O.sql_default_opts(len(rows))))
    for i, r in enumerate(rows):
        ix.add_ref_sequence(100 + i, r.tobytes())
    ix.build()
    res = ix.align_batch(seqs, offs, ids)
    out = []
    for i in range(len(offs) - 1):
        read = seqs[int(offs[i]):int(offs[i + 1])].tobytes()
        ms = []
        for m, rw in zip(ix.matches(res, i, read), res.rows_of(i)):
            d = dict(m.__dict__)
            d["ref_subseq"] = d["ref_subseq"].decode(); d["query_subseq"] = d["query_subseq"].decode()
            d["mapq"] = int(rw["mapq"]); d["nm"] = int(rw["NM"])
            ms.append(d)
        out.append(ms)
    got = make_golden.render(out)
    if got != GOLDEN:
        gl, wl = got.split("\n"), GOLDEN.split("\n")
        diff = [(i, a, b) for i, (a, b) in enumerate(zip(gl, wl)) if a != b][:3]
        raise AssertionError("GPU rows differ from golden (%d vs %d lines): %s" % (len(gl), len(wl), diff))
