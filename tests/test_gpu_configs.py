"""Parity at the sizes the numbers are quoted on: BASELINE configs[0] EXACTLY -- 5 rows x 1 Mbp (seed 20261018),
10 k simulated 150 bp reads (seed 20261019, 0.8 % sub / 0.1 % ins / 0.1 % del), both option sets of BASELINE.md -- and a repeat-planted
5 Mbp reference with 50 k reads, every read compared with the oracle on all PARITY_FIELDS and every CIGAR word."""
import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import synth
from helpers import build_pair, parity_report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("opts_fn", [O.sql_default_opts, O.canonical_opts])
def test_baseline_config0_exact(gpu_lib, opts_fn):
    rows = synth.reference_rows(synth.config_row_lengths("C1"))                       # seed 20261018
    orc, gpu = build_pair(rows, opts_fn(len(rows)))
    seqs, offs, truth = synth.simulate_reads(rows, 10_000, 150)                        # seed 20261019, 1 % error
    ids = synth.lrand48_ids_fast(10_000)
    g = gpu.align_batch(seqs, offs, ids)
    o = orc.align_batch(seqs, offs, ids, 8)
    rep = parity_report(g, o)
    assert rep["reads_checked"] == 10_000 and rep["mismatching_reads"] == 0, rep
    # the simulator's truth as a scale check: the best hit of (nearly) every read is where the read came from
    first = g.row_off[:-1].astype(np.int64)
    pr = g.rows[np.minimum(first, len(g.rows) - 1)]
    ok = (np.diff(g.row_off.astype(np.int64)) > 0) & (pr["rid"] == truth[0]) & (np.abs(pr["pos"] - truth[1]) <= 34)
    assert ok.mean() > 0.995


def test_repeat_planted_5mbp_50k_reads(gpu_lib):
    """SURVEY 8d's realism knob at a size where it bites: 20 repeat families x 8 copies per Mbp (300-6000 bp units, 2 % divergence) in
    5 x 1 Mbp; reads land in repeats => intervals above max_occ are not reached, but re-seeding, chains per read, secondaries, sub / sub_n,
    mem_patch_reg and the warp kernels behind the thread passes all run."""
    rows = synth.reference_rows(synth.config_row_lengths("C1"))
    rows = synth.plant_repeats(rows, n_families=100, copies=8)
    orc, gpu = build_pair(rows, O.sql_default_opts(len(rows)))
    seqs, offs, _ = synth.simulate_reads(rows, 50_000, 150, seed=synth.SEED_READS + 77)
    ids = synth.lrand48_ids_fast(50_000)
    g = gpu.align_batch(seqs, offs, ids)
    o = orc.align_batch(seqs, offs, ids, 8)
    rep = parity_report(g, o)
    assert rep["reads_checked"] == 50_000 and rep["mismatching_reads"] == 0, rep
    assert (g.rows["secondary"] >= 0).sum() > 1000      # the repeats are really hit
    assert int(g.row_off[-1]) > 55_000


def test_repeat_heavy_small_max_occ(gpu_lib):
    """max_occ far below the copy number of a planted family: the k/step sampling rule of mem_chain, frac_rep > 0 and its MAPQ scaling,
    > 32-entry backward lists in the seeding kernel."""
    rows = synth.reference_rows([600_000, 400_001], seed=55)
    rows = synth.plant_repeats(rows, n_families=6, copies=120, unit=(200, 900), divergence=0.01, seed=56)
    opts = O.sql_default_opts(2)
    opts.max_occ = 40
    orc, gpu = build_pair(rows, opts)
    seqs, offs, _ = synth.simulate_reads(rows, 20_000, 150, seed=57)
    ids = synth.lrand48_ids_fast(20_000)
    g = gpu.align_batch(seqs, offs, ids)
    o = orc.align_batch(seqs, offs, ids, 8)
    rep = parity_report(g, o)
    assert rep["mismatching_reads"] == 0, rep
    assert (g.rows["frac_rep"] > 0).sum() > 100


def test_batched_add_ref_equals_loop(gpu_lib):
    """bsq_index_add_ref_datums (texts -> datums on the GPU -> one add call) builds the same index as the add_ref_sequence loop, holes and
    byte-rounding filler included."""
    from bioseqdb_b200 import BwaIndex, _lib
    from helpers import to_bsq
    rows = [r.tobytes() for r in synth.reference_rows([30_001, 1_003, 20_002, 501, 777], seed=61)]
    rows[1] = rows[1][:300] + b"N" * 40 + rows[1][340:600] + b"RRYY" + rows[1][604:]
    rows[3] = b"NN" + rows[3][2:]
    opts = to_bsq(O.sql_default_opts(len(rows)))
    a, b = BwaIndex(0, opts), BwaIndex(0, opts)
    for i, r in enumerate(rows):
        a.add_ref_sequence(10 + i, r)
    b.add_ref_sequences([10 + i for i in range(len(rows))], rows)
    a.build(); b.build()
    for what in range(_lib.ARR_COUNT):
        assert np.array_equal(a.download(what), b.download(what)), what
    assert a.host_state() == b.host_state()
    assert a.n_rows == b.n_rows == len(rows)


def test_replica_with_host_state(gpu_lib):
    """A replica filled from another index's device arrays and host-state blob (what dist.broadcast_index does over NCCL) answers like the
    source: same rows, and the same ref_subseq datums INCLUDING the hole overlay (replicas silently lost their holes)."""
    import ctypes as C
    import torch
    from bioseqdb_b200 import BwaIndex, BsqError, _lib
    from bioseqdb_b200.dist import _CudaArray
    from helpers import to_bsq
    rows = [r.tobytes() for r in synth.reference_rows([50_001, 30_003], seed=71)]
    rows[0] = rows[0][:1000] + b"N" * 25 + rows[0][1025:20_000] + b"SSW" + rows[0][20_003:]
    rows[1] = rows[1][:500] + b"KKKK" + rows[1][504:]
    opts = to_bsq(O.sql_default_opts(2))
    src = BwaIndex(0, opts)
    for i, r in enumerate(rows):
        src.add_ref_sequence(i + 1, r)
    src.build()
    rep = BwaIndex(0, opts)
    m = src.meta()
    _lib.check(rep.L.bsq_index_alloc_replica(rep.h, C.byref(m)))
    for what in range(_lib.ARR_COUNT):
        nb = int(m.arr_bytes[what])
        if not nb:
            continue
        ps, pd = C.c_void_p(), C.c_void_p()
        _lib.check(src.L.bsq_index_device_ptr(src.h, what, C.byref(ps)))
        _lib.check(rep.L.bsq_index_device_ptr(rep.h, what, C.byref(pd)))
        torch.as_tensor(_CudaArray(pd.value, nb), device="cuda").copy_(torch.as_tensor(_CudaArray(ps.value, nb), device="cuda"))
    torch.cuda.synchronize()
    clean = [r.replace(b"N", b"A").replace(b"S", b"C").replace(b"W", b"A").replace(b"K", b"G") for r in rows]
    reads = [rows[0][950:1100].replace(b"N", b"A"), rows[0][19_950:20_100], rows[1][430:580], clean[0][30_000:30_150]]
    from helpers import read_arrays
    seqs, offs = read_arrays(reads)
    ids = synth.lrand48_ids_fast(len(reads))
    rs = src.align_batch(seqs, offs, ids)
    rr = rep.align_batch(seqs, offs, ids)
    with pytest.raises(BsqError, match="replica"):          # an unfinished replica refuses to materialise rows instead of dropping the holes
        rep.tuples(rr, seqs, offs)
    rep.replica_finish(src.host_state())
    assert np.array_equal(rs.row_off, rr.row_off) and len(rs.rows) > 0
    for f in rs.rows.dtype.names:
        if f != "cigar_off":
            assert np.array_equal(rs.rows[f], rr.rows[f]), f
    ts, tr = src.tuples(rs, seqs, offs), rep.tuples(rr, seqs, offs)
    assert np.array_equal(ts.off, tr.off) and np.array_equal(ts.data, tr.data) and np.array_equal(ts.ref_match, tr.ref_match)
    assert [mm.ref_subseq for mm in src.matches(rs, 0, reads[0])] == [mm.ref_subseq for mm in rep.matches(rr, 0, reads[0])]
    assert any(b"N" in mm.ref_subseq for mm in rep.matches(rr, 0, reads[0]))


_DEFAULT_ROUTING = r"""
import sys
sys.path.insert(0, "tests")
import oracle_lib as O
from bioseqdb_b200 import synth
from helpers import build_pair, parity_report
rows = synth.plant_repeats(synth.reference_rows(synth.config_row_lengths("C1")), n_families=100, copies=8)
orc, gpu = build_pair(rows, O.%s(len(rows)))
# noisy reads with long indels: the 9- and 17-column tight passes leave work for the exact-band lists, which are short
seqs, offs, _ = synth.simulate_reads(rows, 60_000, 150, sub=0.03, ins=0.004, dele=0.004, seed=synth.SEED_READS + 91)
ids = synth.lrand48_ids_fast(60_000)
g = gpu.align_batch(seqs, offs, ids)
o = orc.align_batch(seqs, offs, ids, 8)
rep = parity_report(g, o)
assert rep["reads_checked"] == 60_000 and rep["mismatching_reads"] == 0, rep
print("routing ok", int(g.row_off[-1]))
"""


@pytest.mark.parametrize("opts_name", ["sql_default_opts", "canonical_opts"])
def test_default_list_routing(gpu_lib, opts_name):
    """The suite runs with BSQ_SMALL_BATCH_READS=0 and BSQ_FIN_SHORT_LIST=0 so that small cases reach the thread-per-region kernels; the
    library's own defaults -- short job lists handed to the warp-cooperative kernel -- get a 60 k-read case in a process of their own."""
    import os
    import subprocess
    import sys
    env = {k: v for k, v in os.environ.items() if k not in ("BSQ_SMALL_BATCH_READS", "BSQ_FIN_SHORT_LIST")}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", _DEFAULT_ROUTING % opts_name], cwd=root, env=env, capture_output=True, timeout=900)
    assert p.returncode == 0, (p.stdout.decode()[-1500:], p.stderr.decode()[-3000:])
    assert b"routing ok" in p.stdout
