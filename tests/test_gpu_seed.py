"""SMEM intervals of seed_smem vs the oracle's mem_collect_intv, record by record."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import synth, _lib
from helpers import build_pair, read_arrays

pytestmark = pytest.mark.gpu


def _gpu_intervals(gpu, seqs, offs, cap=256):
    n = len(offs) - 1
    out = np.zeros((n, cap, 4), dtype=np.uint64)
    cnt = np.zeros(n, dtype=np.uint32)
    _lib.check(gpu.L.bsq_debug_seed(gpu.h, _lib.ptr(seqs), _lib.ptr(offs), n, _lib.ptr(out), cap, _lib.ptr(cnt)))
    return out, cnt


def test_seed_parity(gpu_lib):
    rows = synth.reference_rows([300_001, 200_002], seed=11)
    rows = synth.plant_repeats(rows, n_families=6, copies=12, unit=(200, 900), divergence=0.01)
    orc, gpu = build_pair(rows, O.sql_default_opts(2))
    seqs, offs, _ = synth.simulate_reads(rows, 600, 150, seed=3, n_frac=0.002)
    out, cnt = _gpu_intervals(gpu, seqs, offs)
    n_ext = 0
    for i in range(600):
        read = seqs[int(offs[i]):int(offs[i + 1])].tobytes()
        iv, _, _ = orc.stage_dump(read)
        assert cnt[i] == len(iv), (i, cnt[i], len(iv))
        assert np.array_equal(out[i, :cnt[i]], iv), i


def test_seed_edge_cases(gpu_lib):
    rows = synth.reference_rows([50_000], seed=12)
    orc, gpu = build_pair(rows, O.sql_default_opts(1))
    ref = rows[0].tobytes()
    reads = [ref[100:250], ref[100:118], ref[100:119], b"N" * 40, ref[500:530] + b"N" + ref[531:600], b"A", ref[1000:1019] + b"NNNN",
             b"ACGT" * 30, ref[49_900:50_000]]
    seqs, offs = read_arrays(reads)
    out, cnt = _gpu_intervals(gpu, seqs, offs)
    for i, r in enumerate(reads):
        iv, _, _ = orc.stage_dump(r)
        assert cnt[i] == len(iv), (i, cnt[i], len(iv))
        assert np.array_equal(out[i, :cnt[i]], iv), i
