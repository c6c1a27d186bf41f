"""the bulk loader.  CPU: the FASTA walk of bioseqdb-import/main.cpp:52-72.  GPU: batch nuclseq_in + nuclseq_from_text
(bsq_nuclseq_from_text_batch) against the host-side codec (bioseqdb_b200/sequence.py, pinned by the payload goldens of SURVEY 8c)."""
import numpy as np
import pytest

from bioseqdb_b200.loader import fasta_records


def test_fasta_records_follow_the_importer():
    data = b">chr1 first\nacgtn\nACGT\n>empty\n>chr2\n\nNNNN\nacgu\n\n>tail"
    assert fasta_records(data) == [(b"chr1 first", b"ACGTNACGT"), (b"chr2", b"NNNNACGU")]
    assert fasta_records(b"ACGT\n>x\nAC\n") == [(b"", b"ACGT"), (b"x", b"AC")]      # sequence before any header: empty name
    assert fasta_records(b"") == []


def test_fasta_stream_and_batches():
    import io
    from bioseqdb_b200.loader import fasta_batches, fasta_record_stream
    data = b">chr1 first\nacgtn\nACGT\n>empty\n>chr2\n\nNNNN\nacgu\n\n>tail"
    assert list(fasta_record_stream(io.BytesIO(data))) == fasta_records(data)
    assert list(fasta_record_stream(io.BytesIO(b"ACGT\n>x\nAC"))) == fasta_records(b"ACGT\n>x\nAC")
    recs = [(b"r%d" % i, b"ACGT" * n) for i, n in enumerate([1, 2, 10, 1, 1, 30, 2])]
    fa = b"".join(b">" + nm + b"\n" + s + b"\n" for nm, s in recs)
    batches = list(fasta_batches(io.BytesIO(fa), batch_bases=40))
    assert [r for b in batches for r in b] == recs                      # order and content kept
    assert all(sum(len(s) for _, s in b) <= 40 or len(b) == 1 for b in batches)
    assert [len(b) for b in batches] == [2, 1, 2, 1, 1]                 # greedy: 4+8 | 40 | 4+4 | 120 (over the cap: alone) | 8
    assert list(fasta_batches(io.BytesIO(b""), 10)) == []


def test_datum_image_layout_matches_the_payload_goldens():
    """SURVEY.md 8c payload goldens (rule of sequence.cpp:209-245) through the datum image the GPU paths are compared with:
    varlena length word << 2, holes_num, len, hole records {i64 offset, i32 len, char amb, 3 pad}, packed codes."""
    import struct
    from bioseqdb_b200.bwa import nuclseq_image
    from bioseqdb_b200.sequence import nuclseq_from_text
    assert nuclseq_image(nuclseq_from_text(b"ACGT")) == struct.pack("<III", 13 << 2, 0, 4) + bytes([0x1b])
    assert nuclseq_image(nuclseq_from_text(b"ACGTA")) == struct.pack("<III", 14 << 2, 0, 5) + bytes([0x1b, 0x39])
    img = nuclseq_image(nuclseq_from_text(b"ACNNGT"))
    assert img == struct.pack("<III", (12 + 16 + 2) << 2, 1, 6) + struct.pack("<qi", 2, 2) + b"N\0\0\0" + bytes([0x16, 0xb9])
    img = nuclseq_image(nuclseq_from_text(b"NNRRA"))
    assert img[:12] == struct.pack("<III", (12 + 32 + 2) << 2, 2, 5)
    assert img[12:44] == struct.pack("<qi", 0, 2) + b"N\0\0\0" + struct.pack("<qi", 2, 2) + b"R\0\0\0" and img[44:] == bytes([0x69, 0x1a])


def _texts(rng, n, max_len):
    out = []
    for i in range(n):
        ln = int(rng.integers(0, max_len + 1)) if i % 7 else int(rng.choice([0, 1, 3, 4, 5, 15, 16, 17, 31, 32, 33, 64]))
        t = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=ln)].copy()
        for _ in range(int(rng.integers(0, 6))):
            if ln == 0:
                break
            a = int(rng.integers(0, ln)); l = int(rng.integers(1, 40))
            t[a:a + l] = ord(rng.choice(list("NNNWSMKRYBDHV")))
        out.append(t.tobytes())
    return out


@pytest.mark.gpu
def test_nuclseq_batch_matches_host_codec(gpu_lib):
    from bioseqdb_b200.bwa import nuclseq_image
    from bioseqdb_b200.loader import nuclseq_images
    from bioseqdb_b200.sequence import nuclseq_from_text
    rng = np.random.default_rng(404)
    texts = _texts(rng, 700, 3000)
    texts += [b"N" * 100_000, b"ACGT" * 25_000 + b"NNR", b"R" * 17 + b"Y" * 16 + b"A" + b"Y" * 15, b"", b"A"]
    images, ms = nuclseq_images(texts)
    assert len(images) == len(texts)
    for i, t in enumerate(texts):
        assert images[i] == nuclseq_image(nuclseq_from_text(t)), (i, len(t))


@pytest.mark.gpu
def test_nuclseq_batch_rejects_what_nuclseq_in_rejects(gpu_lib):
    from bioseqdb_b200._lib import BsqError
    from bioseqdb_b200.loader import nuclseq_images
    with pytest.raises(BsqError, match="invalid nucleotide in nuclseq_in: 'U'"):
        nuclseq_images([b"ACGT", b"ACGUACGT"])
    with pytest.raises(BsqError, match="invalid nucleotide in nuclseq_in: 'a'"):      # lower case is not stored (extension.cpp:40-44)
        nuclseq_images([b"ACGTaCGT"])
    assert nuclseq_images([])[0] == []


@pytest.mark.gpu
def test_load_fasta_end_to_end(gpu_lib):
    from bioseqdb_b200.bwa import nuclseq_image
    from bioseqdb_b200.loader import load_fasta
    from bioseqdb_b200.sequence import nuclseq_from_text
    rows, ms = load_fasta(b">a desc\nacgtnnnn\nACGT\n>b\nNNNNRRYY\nAC\n")
    assert [n for n, _ in rows] == [b"a desc", b"b"]
    assert rows[0][1] == nuclseq_image(nuclseq_from_text(b"ACGTNNNNACGT"))
    assert rows[1][1] == nuclseq_image(nuclseq_from_text(b"NNNNRRYYAC"))
