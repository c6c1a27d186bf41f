"""Index cache: content digest, LRU/eviction and the session id stream -- host logic on CPU with a
stand-in index; the CUDA-backed behaviour is in test_gpu_cache below (gpu marker)."""
import numpy as np
import pytest

from bioseqdb_b200 import synth
from bioseqdb_b200.cache import BwaIndexCache, rows_digest
from bioseqdb_b200.sequence import nuclseq_from_text


class FakeIndex:
    built = 0

    def __init__(self, device):
        self.rows, self.opts, self.is_built, self.closed, self._lrand_state, self.nbytes = [], None, False, False, 0, 100

    def add_ref_sequence(self, rid, seq):
        self.rows.append((rid, seq.len))

    def set_options_from_composite(self, o):
        self.opts = o

    def build(self):
        self.is_built = True
        FakeIndex.built += 1

    def device_bytes(self):
        return self.nbytes

    def close(self):
        self.closed = True

    def align_sequence(self, seq):
        self._lrand_state += 1
        return self._lrand_state


def _rows(seed, n=3):
    return [(i + 1, r.tobytes()) for i, r in enumerate(synth.reference_rows([50 + 7 * i for i in range(n)], seed=seed))]


def test_digest_depends_on_content_ids_and_order():
    rows = [(rid, nuclseq_from_text(t)) for rid, t in _rows(1)]
    d = rows_digest(rows)
    assert d == rows_digest([(rid, nuclseq_from_text(t)) for rid, t in _rows(1)])
    assert d != rows_digest(rows[::-1])
    assert d != rows_digest([(rid + 1, s) for rid, s in rows])
    other = [(rid, nuclseq_from_text(t)) for rid, t in _rows(2)]
    assert d != rows_digest(other)
    # an ambiguity code changes the holes, not the packed bases' length
    t0 = _rows(1)[0][1]
    mod = [(rows[0][0], nuclseq_from_text(t0[:10] + b"N" + t0[11:]))] + rows[1:]
    assert d != rows_digest(mod)


def test_hit_miss_eviction_and_options():
    FakeIndex.built = 0
    c = BwaIndexCache(max_bytes=250, factory=FakeIndex)
    a = c.get(_rows(1), {"min_seed_len": 21})
    assert (c.hits, c.misses, FakeIndex.built) == (0, 1, 1) and a.index.is_built and a.index.opts == {"min_seed_len": 21}
    a2 = c.get(_rows(1), None)
    assert a2.index is a.index and (c.hits, c.misses, FakeIndex.built) == (1, 1, 1)
    assert a2.index.opts is None                      # options are per call, not part of the key
    b = c.get(_rows(2))
    assert b.index is not a.index and FakeIndex.built == 2 and c.evictions == 0
    c.get(_rows(1))                                    # refresh a: b becomes the eviction candidate
    d = c.get(_rows(3))                                # 300 bytes > 250: least recently used goes
    assert c.evictions == 1 and b.index.closed and not a.index.closed and not d.index.closed
    c.get(_rows(2))
    assert FakeIndex.built == 4                        # b had to be rebuilt


def test_session_id_stream_is_shared_between_indexes():
    c = BwaIndexCache(factory=FakeIndex)
    a, b = c.get(_rows(1)), c.get(_rows(2))
    assert a.align_sequence("x") == 1 and b.align_sequence("x") == 2 and a.align_sequence("x") == 3
    assert c.lrand_state == 3


@pytest.mark.gpu
def test_gpu_cache_reuses_resident_index(gpu_lib):
    import oracle_lib as O
    from helpers import compare_results
    from bioseqdb_b200 import BwaIndexCache as Cache
    from bioseqdb_b200 import bwa_opts
    texts = synth.reference_rows([200_003, 100_001], seed=91)
    rows = [(i + 1, t.tobytes()) for i, t in enumerate(texts)]
    cache = Cache(max_bytes=8 << 30)
    a = cache.get(rows, bwa_opts())                    # the composite the SQL function bwa_opts() really delivers
    h0, bytes0 = a.index.h, a.index.device_bytes()
    assert bytes0 > 0
    seqs, offs, _ = synth.simulate_reads(texts, 1500, 120, seed=92)
    r1 = a.align_batch(seqs, offs)                     # ids 0..1499 of the session's lrand48 stream
    a2 = cache.get(rows, bwa_opts())
    assert a2.index.h == h0 and cache.hits == 1 and cache.misses == 1
    r2 = a2.align_batch(seqs, offs)                    # ids 1500..2999: the stream continues across calls
    orc = O.OracleIndex(O.sql_default_opts(len(rows)))
    for rid, t in rows:
        orc.add_ref_text(rid, t)
    orc.build()
    ids = synth.lrand48_ids_fast(3000)
    assert not compare_results(r1, orc.align_batch(seqs, offs, ids[:1500], 4))
    assert not compare_results(r2, orc.align_batch(seqs, offs, ids[1500:], 4))
    assert a2.index.device_bytes() >= bytes0
