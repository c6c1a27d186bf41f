#!/usr/bin/env python
"""Host-side timeline of one chunked bsq_align_batch_datums call (BSQ_TRACE=1): where the e2e time beyond the kernels goes."""
import ctypes as C, os, sys
os.environ["BSQ_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bioseqdb_b200 import BwaIndex, BsqOpts, _lib
from bioseqdb_b200.loader import nuclseq_image_block
args = bench.parse()
rows, seqs, offs, ids, truth = bench.workload(args, 0)
ix = BwaIndex(0, BsqOpts(*bench.opts_tuple(args, len(rows))))
ix.add_ref_sequences(list(range(1, len(rows) + 1)), rows); ix.build()
img, img_off, _ = nuclseq_image_block(seqs, offs, 0)
img_pin = torch.from_numpy(img).pin_memory(); off_pin = torch.from_numpy(img_off.view(np.int64)).pin_memory()
ix.set_rows_ext(False)
resp = C.POINTER(_lib.BsqResult)()
for it in range(4):
    sys.stderr.write("---- call %d\n" % it); sys.stderr.flush()
    ix.session_lrand48(0)
    _lib.check(ix.L.bsq_align_batch_datums(ix.h, C.c_void_p(img_pin.data_ptr()), C.c_void_p(off_pin.data_ptr()), None, args.reads, C.byref(resp)))
    t = ix.timing()
    sys.stderr.write("total %.3f ms (h2d %.3f d2h %.3f; stage sums seed %.2f chain %.2f extend %.2f finalize %.2f)\n" % (t.total, t.h2d, t.d2h, t.seed, t.chain, t.extend, t.finalize))
    ix.L.bsq_result_free(resp)
