"""The C++ host mirror (bioseqdb_b200/host) end to end on the GPU box: the PG-free harness replays the SQL call order, then checks its
own GPU-built tuple columns and bulk-converted datums (8f-4) against the host codec, and its rows must equal the
Python mirror's rows for the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import synth
from helpers import build_pair

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bioseqdb_b200", "host", "harness")


def test_harness_tuples_and_bulk(gpu_lib, tmp_path):
    if not os.path.exists(HARNESS):
        pytest.fail("host harness is not built (run __graft_entry__.build())")
    rng = np.random.default_rng(77)
    rows = []
    for n in (20_001, 10_002):
        t = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
        for _ in range(10):
            a = int(rng.integers(0, n - 40)); l = int(rng.integers(1, 20))
            t[a:a + l] = ord(rng.choice(list("NNRY")))
        rows.append(t)
    seqs, offs, _ = synth.simulate_reads(rows, 300, 150, seed=78)
    seqs = seqs.copy()
    seqs[rng.integers(0, len(seqs), size=100)] = ord("N")
    ref = tmp_path / "ref.tsv"; qry = tmp_path / "q.tsv"
    ref.write_bytes(b"".join(b"%d\t%s\n" % (i + 1, r.tobytes()) for i, r in enumerate(rows)))
    qry.write_bytes(b"".join(b"%d\t%s\n" % (i + 100, seqs[int(offs[i]):int(offs[i + 1])].tobytes()) for i in range(300)))
    p = subprocess.run([HARNESS, str(ref), str(qry), "check_tuples=1", "check_bulk=1"], capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    err = p.stderr.decode()
    assert "check_tuples:" in err and "rows identical" in err and "check_bulk: 2 datums identical" in err
    # the harness rows against the Python mirror (same ids: a fresh process starts lrand48 from its default state)
    orc, gpu = build_pair(rows, O.sql_default_opts(2))
    res = gpu.align_batch(seqs, offs, synth.lrand48_ids_fast(300))
    want = []
    for i in range(300):
        for m in gpu.matches(res, i, seqs[int(offs[i]):int(offs[i + 1])].tobytes()):
            want.append("\t".join(str(x) for x in (m.ref_id, m.ref_subseq.decode(), m.ref_match_begin, m.ref_match_end, m.ref_match_len, i + 100,
                                                   m.query_subseq.decode(), m.query_match_begin, m.query_match_end, m.query_match_len,
                                                   "t" if m.is_primary else "f", "t" if m.is_secondary else "f", "t" if m.is_reverse else "f",
                                                   m.cigar, m.score)))
    assert p.stdout.decode().splitlines() == want
