"""N > 1 host logic on CPU (gloo, world_size 2): contiguous read sharding keeps read i's lrand48 id,
index arrays broadcast from rank 0 arrive intact, per-rank results gather back in read order. The oracle
plays the role of the per-rank aligner (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    from bioseqdb_b200 import synth
    from bioseqdb_b200.dist import broadcast_host_arrays, gather_rows, gather_rows_host, shard_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = synth.reference_rows([60_001, 40_003], seed=91)
    opts = O.sql_default_opts(2)
    seqs, offs, _ = synth.simulate_reads(rows, 301, 150, seed=92)
    ids = synth.lrand48_ids_fast(301)
    ix = O.OracleIndex(opts)
    for i, r in enumerate(rows):
        ix.add_ref_text(i + 1, r.tobytes())
    arrays = None
    if rank == 0:
        ix.build()
        arrays = {"bwt": ix.bwt_plain(), "sa": ix.sa(), "primary": np.array([ix.info()["primary"]], dtype=np.uint64)}
    got = broadcast_host_arrays(arrays, ["bwt", "sa", "primary"], dist)
    if rank != 0:
        ix.adopt(got["bwt"], int(got["primary"][0]), got["sa"])   # replica built from the broadcast arrays only
    s, o, i_, (lo, hi) = shard_reads(seqs, offs, ids, rank, world)
    res = ix.align_batch(s, o, i_, 1)
    merged = gather_rows(res["row_off"], res["rows"], res["cigar"], dist, rank, world)
    merged2 = gather_rows_host(res["row_off"], res["rows"], res["cigar"], dist, None, rank, world)   # byte-tensor gather used by bench.py
    if rank == 0:
        assert all(np.array_equal(a, b) for a, b in zip(merged, merged2))
        full = O.OracleIndex(opts)
        for i, r in enumerate(rows):
            full.add_ref_text(i + 1, r.tobytes())
        full.build()
        ref = full.align_batch(seqs, offs, ids, 1)
        ok = (np.array_equal(merged[0], ref["row_off"]) and np.array_equal(merged[1], ref["rows"]) and np.array_equal(merged[2], ref["cigar"]))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_bounds():
    from bioseqdb_b200.dist import shard_bounds
    for n in [0, 1, 7, 100, 1_000_003]:
        for w in [1, 2, 4, 8]:
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
