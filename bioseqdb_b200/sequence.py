"""Host-side packed nucleotide type: the NUCLSEQ payload of reference bioseqdb/sequence.h:18-38 and its
text codec (bioseqdb/sequence.cpp:46-81,209-245), numpy implementation used by the Python mirror of the
plugin interface. 2 bits per base MSB-first, ambiguity runs kept as (offset, len, letter) "holes", the pac
bits under holes and in the tail padding drawn from std::minstd_rand(holes_num ^ len)."""
from __future__ import annotations

import numpy as np

from ._lib import HOLE_DTYPE

ALLOWED = b"ACGTNWSMKRYBDHV"  # reference sequence.h:16
_NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _NT4[_c] = _i
    _NT4[_c + 32] = _i
_NT4[ord("-")] = 5
_ALLOWED_LUT = np.zeros(256, dtype=bool)
_ALLOWED_LUT[list(ALLOWED)] = True


class NucleotideSequence:
    """len, holes (HOLE_DTYPE array), pac (uint8 array of ceil(len/4) bytes)."""
    __slots__ = ("len", "holes", "pac")

    def __init__(self, length: int, holes: np.ndarray, pac: np.ndarray):
        self.len = int(length)
        self.holes = holes
        self.pac = pac

    @property
    def holes_num(self) -> int:
        return len(self.holes)

    def to_text(self) -> bytes:
        return nuclseq_to_text(self)


def _minstd(seed: int, n: int) -> np.ndarray:
    s = seed % 2147483647
    if s == 0:
        s = 1
    out = np.empty(n, dtype=np.uint32)
    for i in range(n):
        s = s * 48271 % 2147483647
        out[i] = s
    return out


def nuclseq_from_text(text) -> NucleotideSequence:
    """nuclseq_in + nuclseq_from_text (extension.cpp:46-60, sequence.cpp:209-245)."""
    t = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    if len(t) > (2**31 - 1) // 4:
        raise ValueError("provided sequence is too long")
    bad = np.flatnonzero(~_ALLOWED_LUT[t])
    if len(bad):
        raise ValueError("invalid nucleotide in nuclseq_in: '%s'" % chr(int(t[bad[0]])))
    n = len(t)
    codes = _NT4[t]
    amb = codes >= 4
    nbytes = (n + 3) // 4
    vals = np.zeros(nbytes * 4, dtype=np.uint8)
    vals[:n] = codes & 3
    if amb.any() or n % 4:
        # a hole starts where an ambiguous letter differs from the previous character
        prev = np.concatenate([[0], t[:-1]]) if n else t
        starts = np.flatnonzero(amb & (t != prev))
        holes = np.zeros(len(starts), dtype=HOLE_DTYPE)
        amb_idx = np.flatnonzero(amb)
        rng = _minstd(len(starts) ^ n, len(amb_idx) + nbytes * 4 - n)
        vals[amb_idx] = rng[:len(amb_idx)] & 3
        vals[n:] = rng[len(amb_idx):] & 3
        if len(starts):
            # run length: up to the next position that is not the same ambiguous letter
            change = np.flatnonzero(t[1:] != t[:-1]) + 1 if n > 1 else np.array([], dtype=np.int64)
            bounds = np.concatenate([change, [n]])
            ends = bounds[np.searchsorted(bounds, starts, side="right")]
            holes["offset"] = starts
            holes["len"] = ends - starts
            holes["amb"] = t[starts].view("S1")
    else:
        holes = np.zeros(0, dtype=HOLE_DTYPE)
    v = vals.reshape(-1, 4)
    pac = (v[:, 0] << 6 | v[:, 1] << 4 | v[:, 2] << 2 | v[:, 3]).astype(np.uint8)
    return NucleotideSequence(n, holes, pac)


def nuclseq_to_text(s: NucleotideSequence) -> bytes:
    """NucleotideSequence::to_text_palloc / inplace_to_text (sequence.cpp:71-81,162-166)."""
    p = s.pac
    v = np.empty((len(p), 4), dtype=np.uint8)
    v[:, 0] = p >> 6
    v[:, 1] = (p >> 4) & 3
    v[:, 2] = (p >> 2) & 3
    v[:, 3] = p & 3
    out = np.frombuffer(b"ACGT", dtype=np.uint8)[v.reshape(-1)[:s.len]].copy()
    for h in s.holes:
        out[int(h["offset"]):int(h["offset"]) + int(h["len"])] = ord(h["amb"])
    return out.tobytes()
