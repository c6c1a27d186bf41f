#!/usr/bin/env python
"""Latency of small calls through bsq_align_batch with a resident index: what nuclseq_search_bwa (one read per SQL call) sees once the
index is cached.  Prints ms per call for n = 1, 10, 100, 1000, 10000 reads against a 5 x 1 Mbp reference (BASELINE configs[0] shape)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from bioseqdb_b200 import BwaIndex, synth
import oracle_lib as O
from helpers import to_bsq

rows = synth.reference_rows([1_000_000] * 5)
ix = BwaIndex(0, to_bsq(O.sql_default_opts(5)))
for i, r in enumerate(rows):
    ix.add_ref_sequence(i + 1, r.tobytes())
ix.build()
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 10, 100, 1000, 10000]
seqs, offs, _ = synth.simulate_reads(rows, max(sizes), 150, seed=5)
ids = synth.lrand48_ids_fast(max(sizes))
for n in sizes:
    s, o, d = seqs[:int(offs[n])], offs[:n + 1], ids[:n]
    for _ in range(5):
        ix.align_batch(s, o, d)
    reps = 40 if n <= 10000 else 10
    t0 = time.perf_counter()
    for _ in range(reps):
        res = ix.align_batch(s, o, d)
    dt = (time.perf_counter() - t0) / reps
    t = ix.timing()
    print("n=%5d  %.3f ms per call (python wall)   device: seed %.3f chain %.3f extend %.3f finalize %.3f total %.3f ms, %d launches" % (
        n, dt * 1e3, t.seed, t.chain, t.extend, t.finalize, t.total, t.launches))
