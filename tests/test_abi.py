"""The C-ABI library loads and exports every entry point include/bioseqdb_gpu.h declares (no compute
calls: this test runs without a GPU)."""
import os
import re

from bioseqdb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "bioseqdb_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(bsq_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.ABI_SYMBOLS) == declared


def test_struct_layouts():
    import ctypes as C
    assert _lib.ROW_DTYPE.itemsize == 120 and _lib.PUB_ROW_DTYPE.itemsize == 64 and _lib.EXT_ROW_DTYPE.itemsize == 48
    assert C.sizeof(_lib.BsqResult) == 48
    assert C.sizeof(_lib.BsqOpts) == 48
    assert _lib.HOLE_DTYPE.itemsize == 16


def test_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device the constructor must raise."""
    import pytest
    L = _lib.lib()
    if L.bsq_device_count() > 0:
        pytest.skip("a GPU is present")
    from bioseqdb_b200 import BwaIndex, BsqError
    with pytest.raises(BsqError):
        BwaIndex()


def test_bulk_loader_fails_loudly_without_gpu():
    """bsq_nuclseq_from_text_batch needs no index handle: without a CUDA device it must report that, not convert on the host."""
    import pytest
    L = _lib.lib()
    if L.bsq_device_count() > 0:
        pytest.skip("a GPU is present")
    from bioseqdb_b200 import BsqError
    from bioseqdb_b200.loader import nuclseq_images
    with pytest.raises(BsqError, match="no CUDA device"):
        nuclseq_images([b"ACGT"])


def test_new_struct_layouts():
    import ctypes as C
    assert C.sizeof(_lib.BsqTuples) == 48      # n_rows, off, ref_match, bytes, n_bytes, device_ms (+ padding)
    assert C.sizeof(_lib.BsqNuclseqs) == 40    # n, off, bytes, n_bytes, device_ms (+ padding)
