"""row materialisation on the GPU (bsq_result_tuples) against the host-side path the reference takes --
extract_reference_subseq (bwa.cpp:55-68) -> nuclseq_from_text (sequence.cpp:209-245) for ref_subseq, the read's text slice
-> nuclseq_from_text for query_subseq, cigar_compressed_to_string (bwa.cpp:70-77), the int32 ref_match_* (bwa.cpp:171-173)."""
import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import synth
from bioseqdb_b200.bwa import nuclseq_image
from bioseqdb_b200.sequence import nuclseq_from_text
from helpers import build_pair

pytestmark = pytest.mark.gpu


def _check(gpu, rows_text, seqs, offs, res):
    tup = gpu.tuples(res, seqs, offs)
    n_rows = int(res.row_off[-1])
    assert tup.ref_match.shape == (n_rows, 3)
    k = 0
    n_hole_rows = 0
    for i in range(len(offs) - 1):
        text = seqs[int(offs[i]):int(offs[i + 1])].tobytes()
        for m in gpu.matches(res, i, text):
            assert tup.cigar(k) == m.cigar, (i, k)
            assert tup.ref_match[k].tolist() == [m.ref_match_begin, m.ref_match_end, m.ref_match_len], (i, k)
            want_ref = nuclseq_from_text(m.ref_subseq)
            want_q = nuclseq_from_text(m.query_subseq)
            n_hole_rows += (want_ref.holes_num > 0) + (want_q.holes_num > 0)
            assert tup.ref_subseq(k) == nuclseq_image(want_ref), (i, k, m.ref_subseq)
            assert tup.query_subseq(k) == nuclseq_image(want_q), (i, k, m.query_subseq)
            k += 1
    assert k == n_rows
    return n_hole_rows


def test_tuples_plain(gpu_lib):
    rows = synth.reference_rows([60_001, 40_003, 1_001], seed=91)
    orc, gpu = build_pair(rows, O.sql_default_opts(3))
    seqs, offs, _ = synth.simulate_reads(rows, 600, 150, seed=92)
    res = gpu.align_batch(seqs, offs, synth.lrand48_ids_fast(600))
    assert int(res.row_off[-1]) >= 600
    _check(gpu, rows, seqs, offs, res)


def test_tuples_holes_and_ambiguity_codes(gpu_lib):
    """Reference rows with ambiguity runs (un-rebased hole offsets, SURVEY.md B#2: a hole of a later row shadows the same
    offsets of row 1), reads with N / R / Y letters: hole records, the minstd_rand filler under them and in the tail padding."""
    rng = np.random.default_rng(5)
    rows = []
    for n in (30_001, 20_002, 9_999):
        t = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
        for _ in range(12):   # plant ambiguity runs
            a = int(rng.integers(0, n - 40)); l = int(rng.integers(1, 30))
            t[a:a + l] = ord(rng.choice(list("NNRYKM")))
        rows.append(t)
    orc, gpu = build_pair(rows, O.sql_default_opts(3))
    seqs, offs, _ = synth.simulate_reads(rows, 500, 150, seed=93)
    seqs = seqs.copy()
    pick = rng.integers(0, len(seqs), size=400)
    seqs[pick] = np.frombuffer(b"NNNRY", dtype=np.uint8)[rng.integers(0, 5, size=400)]
    for a in rng.integers(0, len(seqs) - 8, size=40):      # runs, and adjacent runs of different letters
        seqs[a:a + 3] = ord("N"); seqs[a + 3:a + 5] = ord("R")
    res = gpu.align_batch(seqs, offs, synth.lrand48_ids_fast(500))
    n_hole_rows = _check(gpu, rows, seqs, offs, res)
    assert n_hole_rows > 20


def test_tuples_fixups(gpu_lib):
    """SURVEY.md 8f-3, opt-in: with hole offsets rebased and reverse-strand hits turned to forward coordinates, ref_subseq is the
    slice [ref_match_begin, ref_match_end) of the reference row exactly as it was inserted -- ambiguity letters included, both strands."""
    from bioseqdb_b200.bwa import TUPLES_FIX_HOLE_OFFSETS, TUPLES_FIX_REVERSE
    rng = np.random.default_rng(17)
    rows = []
    for n in (30_001, 20_002, 9_999):
        t = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
        for _ in range(25):
            a = int(rng.integers(0, n - 40)); l = int(rng.integers(1, 12))
            t[a:a + l] = ord(rng.choice(list("NNRYKM")))
        rows.append(t)
    orc, gpu = build_pair(rows, O.sql_default_opts(3))
    seqs, offs, _ = synth.simulate_reads(rows, 600, 150, seed=95)
    res = gpu.align_batch(seqs, offs, synth.lrand48_ids_fast(600))
    tup = gpu.tuples(res, seqs, offs, TUPLES_FIX_HOLE_OFFSETS | TUPLES_FIX_REVERSE)
    plain = gpu.tuples(res, seqs, offs)
    n_rev = n_amb = 0
    for k, row in enumerate(res.rows):
        b, e, l = tup.ref_match[k].tolist()
        assert l == e - b == int(row["re"] - row["rb"])
        want = rows[int(row["rid"])][b:e].tobytes()
        assert len(want) == l, (k, b, e)
        assert tup.ref_subseq(k) == nuclseq_image(nuclseq_from_text(want)), (k, want)
        assert tup.query_subseq(k) == plain.query_subseq(k) and tup.cigar(k) == plain.cigar(k)
        n_rev += int(row["is_rev"]); n_amb += any(c not in b"ACGT" for c in want)
        assert b == int(row["pos"])          # mem_aln_t.pos: row-relative, forward strand, leftmost base
    assert n_rev > 100 and n_amb > 10


def test_tuples_empty(gpu_lib):
    rows = synth.reference_rows([20_001], seed=94)
    orc, gpu = build_pair(rows, O.sql_default_opts(1))
    seqs = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)     # shorter than min_seed_len: no rows
    offs = np.array([0, 10], dtype=np.uint64)
    res = gpu.align_batch(seqs, offs, synth.lrand48_ids_fast(1))
    assert int(res.row_off[-1]) == 0
    tup = gpu.tuples(res, seqs, offs)
    assert tup.off.tolist() == [0] and tup.ref_match.shape == (0, 3)
