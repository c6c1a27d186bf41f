"""Deterministic synthetic references and simulated reads (SURVEY.md 8d / BASELINE.md section 3).

Generator: numpy PCG64 streams seeded 20261018 (reference), 20261019 (reads), 20261020 (repeat
planting).  References are i.i.d. uniform ACGT rows ``(id, text)`` with ids 1..rows; reads start
uniformly over (row, offset) fully inside one row, strand 50/50, with per-base substitution /
insertion / deletion errors.  Everything is host-side numpy: this module feeds both the GPU path
and the CPU baseline with identical bytes.
"""
from __future__ import annotations

import numpy as np

SEED_REF = 20261018
SEED_READS = 20261019
SEED_REPEATS = 20261020

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b

# human chromosome lengths (Mbp, GRCh38 1..22, X, Y) -- only their proportions are used (config C3)
_HUMAN_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]


def reference_rows(row_lengths, seed: int = SEED_REF):
    """Return a list of uint8 ASCII arrays, one per reference row."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [_ACGT[rng.integers(0, 4, size=int(n), dtype=np.uint8)] for n in row_lengths]


def config_row_lengths(config: str):
    if config == "C1":
        return [1_000_000] * 5
    if config == "C2":
        return [10_000_000] * 10
    if config == "C3":
        tot = sum(_HUMAN_MBP)
        return [int(3_100_000_000 * m / tot) for m in _HUMAN_MBP]
    if config == "C5":
        rng = np.random.Generator(np.random.PCG64(SEED_REF + 5))
        ln = rng.integers(500, 1501, size=500_000)
        ln = ln + (ln % 4 == 0)  # lengths not multiples of 4: byte-rounding filler is exercised
        return ln.tolist()
    raise ValueError(config)


def plant_repeats(rows, n_families=20, copies=8, unit=(300, 6000), divergence=0.02, seed: int = SEED_REPEATS):
    """Optional realism knob: overwrite random places with diverged copies of repeat units."""
    rng = np.random.Generator(np.random.PCG64(seed))
    for _ in range(n_families):
        ulen = int(rng.integers(unit[0], unit[1] + 1))
        u = _ACGT[rng.integers(0, 4, size=ulen, dtype=np.uint8)]
        for _ in range(copies):
            r = int(rng.integers(0, len(rows)))
            if len(rows[r]) <= ulen:
                continue
            p = int(rng.integers(0, len(rows[r]) - ulen))
            c = u.copy()
            mut = rng.random(ulen) < divergence
            c[mut] = _ACGT[rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)]
            rows[r][p:p + ulen] = c
    return rows


def simulate_reads(rows, n_reads: int, read_len: int = 150, sub=0.008, ins=0.001, dele=0.001,
                   seed: int = SEED_READS, n_frac: float = 0.0, chunk: int = 200_000):
    """Simulate reads. Returns (seqs uint8[n_reads*read_len] ASCII, offs uint64[n+1], truth) where
    truth = (row index, offset, strand) arrays kept for sanity only."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([len(r) for r in rows], dtype=np.int64)
    slack = max(8, int(read_len * (dele * 4 + 0.02)) + 8)
    tpl = read_len + slack
    ok = lens >= tpl
    if not ok.any():
        raise ValueError("rows shorter than read template")
    weights = np.where(ok, lens - tpl + 1, 0).astype(np.float64)
    cum = np.cumsum(weights)
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    t_row = np.empty(n_reads, dtype=np.int64)
    t_off = np.empty(n_reads, dtype=np.int64)
    t_rev = np.empty(n_reads, dtype=np.bool_)
    # concatenated view for fancy indexing
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]
    cat = np.concatenate(rows) if len(rows) > 1 else rows[0]
    for lo in range(0, n_reads, chunk):
        n = min(chunk, n_reads - lo)
        u = rng.random(n) * cum[-1]
        ri = np.searchsorted(cum, u, side="right")
        ri = np.minimum(ri, len(rows) - 1)
        off = (rng.random(n) * weights[ri]).astype(np.int64)
        off = np.minimum(off, (weights[ri] - 1).astype(np.int64))
        rev = rng.random(n) < 0.5
        idx = (starts[ri] + off)[:, None] + np.arange(tpl, dtype=np.int64)[None, :]
        t = cat[idx]  # n x tpl template, forward strand
        # reverse strand: reverse-complement the template window
        t[rev] = _COMP[t[rev][:, ::-1]]
        # per-template-base edits
        r = rng.random((n, tpl))
        is_del = r < dele
        is_ins = (r >= dele) & (r < dele + ins)
        is_sub = (r >= dele + ins) & (r < dele + ins + sub)
        nsub = int(is_sub.sum())
        if nsub:
            cur = t[is_sub]
            code = np.searchsorted(_ACGT, cur)  # ACGT is sorted in ASCII
            t[is_sub] = _ACGT[(code + rng.integers(1, 4, size=nsub)) % 4]
        ins_base = _ACGT[rng.integers(0, 4, size=(n, tpl), dtype=np.uint8)]
        # two slots per template base: [base unless deleted][inserted base if any]
        slots = np.empty((n, tpl * 2), dtype=np.uint8)
        slots[:, 0::2] = t
        slots[:, 1::2] = ins_base
        valid = np.empty((n, tpl * 2), dtype=np.bool_)
        valid[:, 0::2] = ~is_del
        valid[:, 1::2] = is_ins
        rank = np.cumsum(valid, axis=1) - 1
        take = valid & (rank < read_len)
        rows_i, cols_i = np.nonzero(take)
        o = np.zeros((n, read_len), dtype=np.uint8)
        o[rows_i, rank[rows_i, cols_i]] = slots[rows_i, cols_i]
        assert (o != 0).all(), "template slack too small"
        if n_frac > 0:
            o[rng.random((n, read_len)) < n_frac] = ord("N")
        out[lo:lo + n] = o
        t_row[lo:lo + n] = ri
        t_off[lo:lo + n] = np.where(rev, off + (tpl - read_len), off)  # error-free reverse reads end at the window end
        t_rev[lo:lo + n] = rev
    offs = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return out.reshape(-1), offs, (t_row, t_off, t_rev)


def lrand48_ids(n: int, start_state: int = 0):
    """The ids mem_align1 would draw: the first n outputs of glibc lrand48() from a fresh process
    (default state 0; SURVEY.md A.10 / 8c golden 0, 2116118, 89401895, ...). Returns (ids, state)."""
    a, c, m = 0x5DEECE66D, 0xB, (1 << 48) - 1
    out = np.empty(n, dtype=np.int64)
    x = start_state
    for i in range(n):
        x = (x * a + c) & m
        out[i] = x >> 17
    return out, x


def lrand48_ids_fast(n: int):
    """Vectorised lrand48 stream (jump-ahead by doubling) for large n."""
    a, c, m = 0x5DEECE66D, 0xB, (1 << 48) - 1
    if n <= 4096:
        return lrand48_ids(n)[0]
    # x_{i+k} = A_k x_i + C_k ; build by blocks of 4096 with python ints for the block heads
    blk = 4096
    A, C = 1, 0
    As = np.empty(blk, dtype=object); Cs = np.empty(blk, dtype=object)
    for k in range(blk):
        A = (A * a) & m; C = (C * a + c) & m
        As[k] = A; Cs[k] = C
    out = np.empty(n, dtype=np.int64)
    x = 0
    As_l = [int(v) for v in As]; Cs_l = [int(v) for v in Cs]
    # split 48-bit multiply into numpy uint64-safe pieces: use python ints per block head only,
    # and per element (A_k * x + C_k) mod 2^48 via 24-bit limbs
    A_lo = np.array([v & 0xFFFFFF for v in As_l], dtype=np.uint64)
    A_hi = np.array([v >> 24 for v in As_l], dtype=np.uint64)
    Cv = np.array(Cs_l, dtype=np.uint64)
    M = np.uint64(m)
    for lo in range(0, n, blk):
        k = min(blk, n - lo)
        x_lo = np.uint64(x & 0xFFFFFF); x_hi = np.uint64(x >> 24)
        # (A_hi*2^24 + A_lo)(x_hi*2^24 + x_lo) mod 2^48 = A_lo*x_lo + ((A_hi*x_lo + A_lo*x_hi) mod 2^24) * 2^24
        low = A_lo[:k] * x_lo
        mid = ((A_hi[:k] * x_lo + A_lo[:k] * x_hi) & np.uint64(0xFFFFFF)) << np.uint64(24)
        v = (low + mid + Cv[:k]) & M
        out[lo:lo + k] = (v >> np.uint64(17)).astype(np.int64)
        x = int(v[k - 1])
    return out
