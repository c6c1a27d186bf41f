"""The claim behind the tight first tries of the finalize stage (bioseqdb_b200/csrc/finalize.cu, regs_cigar_narrow<4> / <3>; DESIGN.md
2.5), checked on the CPU oracle's ksw_global2 alone: if the banded global alignment inside a window of half-width w' scores MORE than any
path that leaves the window possibly can -- ub = a (len - k) - gapcost(k) - gapcost(k -+ (lq - rlen)), k = w' + 1 -- then score AND
CIGAR equal those of the band bwa_gen_cigar2 would have asked for, whatever that band is.  Low-complexity sequences (gap placement ties
everywhere) are part of the sample; the test also counts how often the premise holds so that it cannot pass vacuously."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O


def _ub(opts, lq, rlen, wt):
    # the largest matrix entry is 1 whatever `a` says: the path's matrix is mem_opt_init's and is never refilled;
    # the kernels read it from the matrix itself (DevOpts::mat_max)
    k, dl, a = wt + 1, lq - rlen, 1
    ubp = a * (lq - k) - (opts.o_ins + opts.e_ins * k) - (opts.o_del + opts.e_del * (k - dl))
    ubm = a * (rlen - k) - (opts.o_del + opts.e_del * k) - (opts.o_ins + opts.e_ins * (k + dl))
    return max(ubp, ubm)


def _global(L, q, t, opts, w):
    cig = np.zeros(1024, dtype=np.uint32)
    n = C.c_int()
    sc = L.orc_ksw_global2(len(q), O._ptr(q), len(t), O._ptr(t), C.byref(opts), w, O._ptr(cig), 1024, C.byref(n))
    return sc, cig[:n.value].tolist()


def _pair(rng, style):
    """(query, target) as nt4 arrays: a target of 60..190 bases and a query derived from it with substitutions and indels."""
    n = int(rng.integers(60, 190))
    if style == 0:
        t = rng.integers(0, 4, size=n)
    elif style == 1:      # homopolymer runs and short tandem repeats: ties between gap positions
        t = np.repeat(rng.integers(0, 4, size=n // 3 + 1), rng.integers(1, 7, size=n // 3 + 1))[:n]
    else:
        unit = rng.integers(0, 4, size=int(rng.integers(2, 5)))
        t = np.tile(unit, n // len(unit) + 1)[:n]
        t = np.where(rng.random(n) < 0.05, rng.integers(0, 4, size=n), t)
    q = []
    sub, indel = rng.choice([0.005, 0.02, 0.06]), rng.choice([0.002, 0.01, 0.03])
    for b in t:
        r = rng.random()
        if r < indel:
            continue                                  # deletion from the query
        if r < 2 * indel:
            q.extend(rng.integers(0, 4, size=int(rng.integers(1, 4))).tolist())   # insertion
        q.append(int(b) if rng.random() >= sub else int(rng.integers(0, 4)))
    return np.array(q, dtype=np.uint8), t.astype(np.uint8)


@pytest.mark.parametrize("opts_fn", [O.sql_default_opts, O.canonical_opts,
                                     lambda n: O.Opts(19, 500, 2, 3, 5, 5, 100, 100, 4, 2, 5, 1)])
def test_tight_band_result_is_the_full_band_result(oracle, opts_fn):
    L = O.lib()
    opts = opts_fn(1)
    rng = np.random.default_rng(20261019)
    held = {4: 0, 8: 0}
    tried = 0
    for it in range(2500):
        q, t = _pair(rng, it % 3)
        lq, rlen = len(q), len(t)
        if lq == 0 or abs(lq - rlen) > 8:
            continue
        for w_full in (int(rng.integers(abs(lq - rlen) + 9, 60)), 100):
            full = None
            for wt in (4, 8):
                if abs(lq - rlen) > wt or wt >= w_full:
                    continue
                tried += 1
                sc, cig = _global(L, q, t, opts, wt)
                if sc > _ub(opts, lq, rlen, wt):
                    if full is None:
                        full = _global(L, q, t, opts, w_full)
                    assert (sc, cig) == full, (it, wt, w_full, lq, rlen)
                    held[wt] += 1
    print("tight-band claim: windows tried", tried, "premise held", held)
    assert tried > 4000 and held[4] > 0.3 * tried / 2 and held[8] > held[4] * 0.9, (tried, held)


@pytest.mark.parametrize("opts_fn", [O.sql_default_opts, O.canonical_opts])
def test_diagonal_proof_of_the_equal_length_pass(oracle, opts_fn):
    """Pass 0 of the finalize stage (regs_cigar_narrow<5>): for a region with lq == rlen, a gap-free diagonal scoring more than
    (lq - 1) max(mat) - oe_ins - oe_del is the unique optimum of ksw_global2 under EVERY band -- CIGAR lq M, score = the diagonal's."""
    L = O.lib()
    opts = opts_fn(1)
    rng = np.random.default_rng(7)
    held = 0
    for it in range(1500):
        n = int(rng.integers(20, 180))
        t = rng.integers(0, 4, size=n).astype(np.uint8) if it % 2 else np.repeat(rng.integers(0, 4, size=n // 2 + 1), rng.integers(1, 5, size=n // 2 + 1))[:n].astype(np.uint8)
        n = len(t)
        q = np.where(rng.random(n) < rng.choice([0.0, 0.01, 0.03]), rng.integers(0, 4, size=n), t).astype(np.uint8)
        diag = int(np.sum(np.where(q == t, 1, -4)))
        if diag <= (n - 1) * 1 - (opts.o_ins + opts.e_ins) - (opts.o_del + opts.e_del):
            continue
        held += 1
        for w in (1, 5, 37, 100):
            sc, cig = _global(L, q, t, opts, w)
            assert sc == diag and cig == [n << 4], (it, w, n, diag, sc, cig)
    assert held > 700, held
