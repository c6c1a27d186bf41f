#!/usr/bin/env python
"""Queue sizes of the finalize passes on a bench.py workload (default c2): how many regions each narrow-band kernel sees.
Usage: python scripts/fin_queues.py [bench.py flags]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from bioseqdb_b200 import BwaIndex, BsqOpts, _lib

args = bench.parse()
rows, seqs, offs, ids, truth = bench.workload(args, 0)
ix = BwaIndex(0, BsqOpts(*bench.opts_tuple(args, len(rows))))
ix.add_ref_sequences(list(range(1, len(rows) + 1)), rows)
ix.build()
ix.upload(seqs, offs, ids)
ix.set_counters(True)
ix.align_resident(); ix.align_resident()
ctl = np.zeros(64, np.uint32)
_lib.check(ix.L.bsq_debug_ctl(ix.h, ctl.ctypes.data_as(C.c_void_p)))
names = ["A front (DP tries 1)", "B front (tries 2)", "A front (tries 3)", "S (equal-length, diagonal pass)", "sink", "A back", "B back", "A back (tries 3)"]
for k, nme in enumerate(names):
    print("%-34s %9d" % (nme, ctl[32 + k]))
print("%-34s %9d" % ("T4 (tight first tries, 9 columns)", ctl[52]))
print("%-34s %9d" % ("T8 (17 columns)", ctl[53]))
print("reads left for regs_finalize (warp per read)", ctl[58])
print("wide jobs", ctl[25], "cigar words", ctl[6], "counters", ix.counters())
res = ix.download_result()
print("rows", int(res.row_off[-1]), "digest", bench.rows_digest(res))
t = ix.timing()
print("stages ms: seed %.2f chain %.2f extend %.2f finalize %.2f total %.2f" % (t.seed, t.chain, t.extend, t.finalize, t.total))
