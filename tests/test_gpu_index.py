"""the GPU-built FM-index (text, L2, BWT, primary, Occ checkpoints, SA) must
equal the oracle's, which is pinned by brute-force suffix sorting (test_oracle.py)."""
import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import synth
from bioseqdb_b200 import _lib
from helpers import build_pair

pytestmark = pytest.mark.gpu


def _check_index(orc, gpu):
    oi, m = orc.info(), gpu.meta()
    assert m.built == 1
    assert (m.l_pac, m.seq_len, m.primary) == (oi["l_pac"], oi["seq_len"], oi["primary"])
    assert list(m.L2) == oi["L2"]
    assert np.array_equal(gpu.bwt_plain(), orc.bwt_plain())
    assert np.array_equal(gpu.sa_sampled(), orc.sa())
    assert np.array_equal(gpu.download(_lib.ARR_PAC), orc.pac())
    # Occ checkpoints: GPU keeps 64-byte blocks {4 x u64 counts, 8 x u32}; the oracle has bwa's layout,
    # identical for every full block
    g = gpu.download(_lib.ARR_OCC).view(np.uint32)
    o = orc.bwt_interleaved()
    n_full = oi["seq_len"] // 128
    assert np.array_equal(g[:n_full * 16], o[:n_full * 16])
    # trailing count record = L2 differences
    n_blocks = (oi["seq_len"] + 127) // 128
    tail = g[n_blocks * 16:n_blocks * 16 + 8].view(np.uint64)
    assert tail.tolist() == [oi["L2"][c + 1] - oi["L2"][c] for c in range(4)]
    # full SA against the oracle's bwt_sa on a sample of rows
    sa = gpu.download(_lib.ARR_SA).view(np.uint32 if m.sa_bytes == 4 else np.uint64)
    assert int(sa[0]) == oi["seq_len"]
    rng = np.random.default_rng(5)
    for k in rng.integers(1, oi["seq_len"] + 1, size=300):
        assert int(sa[int(k)]) == orc.bwt_sa(int(k))
    # the device-side verifier (no code shared with the builder) agrees: every row and block, nothing undecided on these texts
    chk = gpu.verify(n_samples=1 << 40)
    assert chk["sound"] and chk["exhaustive"] == 1 and chk["rows_checked"] == oi["seq_len"] + 1 and chk["order_checked"] == oi["seq_len"], chk


@pytest.mark.parametrize("lens", [[1000], [997, 1503, 64], [200_003, 150_001, 99_999]])
def test_index_random(gpu_lib, lens):
    rows = synth.reference_rows(lens, seed=synth.SEED_REF + len(lens))
    orc, gpu = build_pair(rows, O.sql_default_opts(len(rows)))
    _check_index(orc, gpu)


@pytest.mark.parametrize("text", [b"A" * 300, b"ACGT" * 100, b"AC" * 257 + b"G", b"ACGTTGCA" * 40 + b"A" * 100, b"T", b"ACG"])
def test_index_repetitive(gpu_lib, text):
    """Low-complexity texts need many prefix-doubling rounds (the initial 28-mer keys tie)."""
    orc, gpu = build_pair([text], O.sql_default_opts(1))
    _check_index(orc, gpu)


def test_index_with_holes(gpu_lib):
    rows = [b"ACGTNNNNACGTRYACGT" * 30 + b"AC", b"N" * 50 + b"ACGT" * 50]
    orc, gpu = build_pair(rows, O.sql_default_opts(2))
    _check_index(orc, gpu)


def test_empty_reference(gpu_lib):
    """bwa.cpp:108-109,142-143: no rows => build is a no-op and alignment returns nothing."""
    from bioseqdb_b200 import BwaIndex
    ix = BwaIndex()
    ix.build()
    assert ix.align_sequence(b"ACGTACGTACGTACGTACGTACGT") == []


@pytest.mark.parametrize("text", [None, b"A" * 300, b"ACGTTGCA" * 40 + b"A" * 100, b"AC" * 257 + b"G"])
def test_index_wide_path_small(gpu_lib, monkeypatch, text):
    """The 64-bit-id build (bucketed counting sort + per-bucket radix sort + compacted prefix doubling) and the
    64-bit seeding path, forced on small inputs so that the oracle can check them (at >= 2^32 rows it cannot)."""
    monkeypatch.setenv("BSQ_FORCE_WIDE", "1")
    rows = [text] if text is not None else synth.reference_rows([120_001, 80_003, 997], seed=55)
    orc, gpu = build_pair(rows, O.sql_default_opts(len(rows)))
    assert gpu.meta().sa_bytes == 8
    _check_index(orc, gpu)
    if text is None:
        from helpers import compare_results
        seqs, offs, _ = synth.simulate_reads(rows, 800, 150, seed=56)
        ids = synth.lrand48_ids_fast(800)
        bad = compare_results(gpu.align_batch(seqs, offs, ids), orc.align_batch(seqs, offs, ids, 2))
        assert not bad, "\n".join(bad)


def _device_view(gpu, what, dtype):
    """torch view of one of the index's device arrays (the handle the multi-GPU broadcast writes through)."""
    import ctypes as C
    import torch
    from bioseqdb_b200.dist import _CudaArray
    p = C.c_void_p()
    _lib.check(gpu.L.bsq_index_device_ptr(gpu.h, what, C.byref(p)))
    n = int(gpu.meta().arr_bytes[what])
    return torch.as_tensor(_CudaArray(p.value, n), device="cuda").view(dtype)


@pytest.mark.parametrize("wide", [False, True])
def test_index_verifier_catches_corruption(gpu_lib, monkeypatch, wide):
    """bsq_index_verify is the check the 6.2 G-row index gets (nothing else can look at it): here it must pass a sound index and name
    each kind of damage -- two suffix-array rows swapped, one BWT symbol changed, one Occ checkpoint off by one."""
    import torch
    if wide:
        monkeypatch.setenv("BSQ_FORCE_WIDE", "1")
    rows = synth.reference_rows([50_001, 30_003], seed=77)
    _, gpu = build_pair(rows, O.sql_default_opts(2))
    assert gpu.meta().sa_bytes == (8 if wide else 4)
    ok = gpu.verify(n_samples=1 << 40)
    assert ok["sound"] and ok["order_undecided"] == 0, ok
    sampled = gpu.verify(n_samples=5000)
    assert sampled["sound"] and sampled["exhaustive"] == 0 and sampled["rows_checked"] == 5000, sampled
    sa = _device_view(gpu, _lib.ARR_SA, torch.int64 if wide else torch.int32)
    occ = _device_view(gpu, _lib.ARR_OCC, torch.int32)
    # (1) two rows of the suffix array swapped: still a permutation, no longer sorted, LF and BWT break at those rows
    a, b = int(sa[1000]), int(sa[1001])
    sa[1000], sa[1001] = b, a
    torch.cuda.synchronize()
    bad = gpu.verify(n_samples=1 << 40)
    assert not bad["sound"] and bad["sa_permutation_ok"] == 1 and bad["order_bad"] >= 1 and bad["lf_bad"] >= 1, bad
    sa[1000], sa[1001] = a, b
    # (2) a duplicated entry: not a permutation any more
    sa[2000] = int(sa[2001])
    torch.cuda.synchronize()
    bad = gpu.verify(n_samples=1 << 40)
    assert not bad["sound"] and bad["sa_permutation_ok"] == 0, bad
    # restore the entry from a second build of the same rows
    _, fresh = build_pair(rows, O.sql_default_opts(2))
    good_sa = fresh.download(_lib.ARR_SA).view(np.uint64 if wide else np.uint32)
    sa[2000] = int(good_sa[2000])
    torch.cuda.synchronize()
    assert gpu.verify(n_samples=1 << 40)["sound"]
    # (3) one BWT symbol changed (block 3, first word): the BWT check, LF and the block recount notice
    w = int(occ[3 * 16 + 8])
    occ[3 * 16 + 8] = w ^ (1 << 30)
    torch.cuda.synchronize()
    bad = gpu.verify(n_samples=1 << 40)
    assert not bad["sound"] and bad["bwt_bad"] >= 1 and bad["occ_bad"] >= 1, bad
    occ[3 * 16 + 8] = w
    # (4) one checkpoint count off by one
    c = int(occ[5 * 16])
    occ[5 * 16] = c + 1
    torch.cuda.synchronize()
    bad = gpu.verify(n_samples=1 << 40)
    assert not bad["sound"] and bad["occ_bad"] >= 1 and bad["lf_bad"] >= 1, bad
    occ[5 * 16] = c
    torch.cuda.synchronize()
    assert gpu.verify(n_samples=1 << 40)["sound"]
