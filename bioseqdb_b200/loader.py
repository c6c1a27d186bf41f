"""Bulk loader: the FASTA walk of reference bioseqdb-import/main.cpp:52-72 on the host, then ONE call that turns
every record's text into a finished NUCLSEQ datum on the GPU (bsq_nuclseq_from_text_batch = nuclseq_in + nuclseq_from_text,
extension.cpp:46-60, sequence.cpp:209-245) -- instead of one INSERT and one server-side conversion per record."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BsqNuclseqs, check, ptr


def fasta_records(data: bytes):
    """(name, sequence) pairs exactly as bioseqdb-import submits them (main.cpp:52-72): std::getline lines, a line starting with
    '>' opens a record named by the rest of the line, other lines are upper-cased and appended, records without sequence are dropped."""
    name, seq, out = b"", [], []

    def submit():
        nonlocal name, seq
        s = b"".join(seq)
        if s:
            out.append((name, s))
        name, seq = b"", []

    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                       # getline does not produce an empty line after the final newline
    for line in lines:
        if line[:1] == b">":
            submit()
            name = line[1:]
        else:
            seq.append(line.upper())
    submit()
    return out


def fasta_record_stream(lines):
    """The same walk as fasta_records over an iterable of lines (bytes, with or without the trailing newline): yields (name, sequence)
    one record at a time, so that a file of any size is read in bounded memory."""
    name, seq = b"", []
    for line in lines:
        if line.endswith(b"\n"):
            line = line[:-1]
        if line[:1] == b">":
            s = b"".join(seq)
            if s:
                yield name, s
            name, seq = line[1:], []
        else:
            seq.append(line.upper())
    s = b"".join(seq)
    if s:
        yield name, s


def fasta_batches(lines, batch_bases: int = 1 << 30):
    """Groups the records of fasta_record_stream into batches of at most batch_bases bases (a record longer than that is a batch of its
    own): one bsq_nuclseq_from_text_batch call per batch keeps the device buffers bounded (text + images ~ 1.3 bytes per base)."""
    batch, n = [], 0
    for rec in fasta_record_stream(lines):
        if batch and n + len(rec[1]) > batch_bases:
            yield batch
            batch, n = [], 0
        batch.append(rec)
        n += len(rec[1])
    if batch:
        yield batch


def nuclseq_image_block(cat: np.ndarray, offs: np.ndarray, device: int = 0):
    """n sequences (concatenated upper-case text `cat`, offs[n + 1]) -> (bytes, off[n + 1], device_ms): the datum images back to back as
    the library wrote them (8-byte aligned; the true size of an image is in its length word), ready for bsq_align_batch_datums /
    bsq_index_add_ref_datums.  off[n] = total bytes."""
    L = _lib.lib()
    cat = np.ascontiguousarray(cat, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    res = C.POINTER(BsqNuclseqs)()
    check(L.bsq_nuclseq_from_text_batch(device, ptr(cat if len(cat) else np.zeros(1, dtype=np.uint8)), ptr(offs), n, C.byref(res)))
    r = res.contents
    off = np.ctypeslib.as_array(r.off, shape=(n + 1,)).copy()
    nb = int(r.n_bytes)
    data = np.ctypeslib.as_array(r.bytes, shape=(max(nb, 1),))[:nb].copy()
    ms = float(r.device_ms)
    L.bsq_nuclseqs_free(res)
    return data, off, ms


def nuclseq_images(texts, device: int = 0):
    """texts: list of bytes. Returns (list of datum images, device_ms). Raises BsqError like nuclseq_in on an invalid letter."""
    L = _lib.lib()
    n = len(texts)
    offs = np.zeros(n + 1, dtype=np.uint64)
    if n:
        offs[1:] = np.cumsum([len(t) for t in texts])
    cat = np.frombuffer(b"".join(texts), dtype=np.uint8) if n and offs[-1] else np.zeros(1, dtype=np.uint8)
    res = C.POINTER(BsqNuclseqs)()
    check(L.bsq_nuclseq_from_text_batch(device, ptr(cat), ptr(offs), n, C.byref(res)))
    r = res.contents
    off = np.ctypeslib.as_array(r.off, shape=(n + 1,)).copy()
    nb = int(r.n_bytes)
    data = np.ctypeslib.as_array(r.bytes, shape=(max(nb, 1),))[:nb].copy()
    ms = float(r.device_ms)
    L.bsq_nuclseqs_free(res)
    images = []
    for i in range(n):
        at = int(off[i])
        size = int(np.frombuffer(data[at:at + 4].tobytes(), dtype="<u4")[0]) >> 2
        images.append(data[at:at + size].tobytes())
    return images, ms


def load_fasta_file(path: str, device: int = 0, batch_bases: int = 1 << 30):
    """Yields (name, NUCLSEQ datum image) for every record of a FASTA file, converting batch_bases bases per GPU call."""
    with open(path, "rb") as f:
        for batch in fasta_batches(f, batch_bases):
            images, _ = nuclseq_images([s for _, s in batch], device)
            for (nm, _), img in zip(batch, images):
                yield nm, img


def load_fasta(data: bytes, device: int = 0):
    """[(name, NUCLSEQ datum image)] for a FASTA file's bytes -- the rows a COPY ... BINARY into (name, seq) would carry."""
    recs = fasta_records(data)
    images, ms = nuclseq_images([s for _, s in recs], device)
    return [(nm, img) for (nm, _), img in zip(recs, images)], ms
