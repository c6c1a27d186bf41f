"""Library-level multi-GPU (reference extension.cpp:346-377 is one process, one thread): bsq_multi_* replicates a built
index to a second device with peer copies and cuts a batch over both devices from ONE host thread; the rows must equal the
single-device rows, and the oracle's.  Needs two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import MultiBwaIndex, synth
from bioseqdb_b200.loader import nuclseq_image_block
from helpers import build_pair, compare_results, PARITY_FIELDS

pytestmark = pytest.mark.gpu


def test_two_devices_equal_one(gpu_lib):
    if gpu_lib.bsq_device_count() < 2:
        pytest.skip("needs two GPUs")
    rows = [r.tobytes() for r in synth.reference_rows([300_001, 200_003], seed=151)]
    rows[1] = rows[1][:7000] + b"NNNNNNNN" + rows[1][7008:]
    orc, gpu = build_pair(rows, O.sql_default_opts(2))
    clean = [np.frombuffer(r.replace(b"N", b"C"), dtype=np.uint8).copy() for r in rows]
    clean = synth.plant_repeats(clean, n_families=4, copies=5, unit=(150, 300), divergence=0.0, seed=152)   # equal scores: hash(id) decides the order
    n = 30_011
    seqs, offs, _ = synth.simulate_reads(clean, n, 150, seed=153, n_frac=0.002)
    ids = synth.lrand48_ids_fast(n)
    one = gpu.align_batch(seqs, offs, ids)
    multi = MultiBwaIndex(gpu, [0, 1])
    two = multi.align_batch(seqs, offs, ids)
    assert np.array_equal(one.row_off, two.row_off)
    for f in PARITY_FIELDS:
        assert np.array_equal(one.rows[f], two.rows[f]), f
    for i in range(0, len(one.rows), 37):
        assert one.cigar_of(one.rows[i]) == two.cigar_of(two.rows[i])
    assert not compare_results(two, orc.align_batch(seqs, offs, ids, 8))
    # datum images + ids from the session stream, split over the devices: read i keeps the i-th draw
    data, off, _ = nuclseq_image_block(seqs, offs)
    gpu.session_lrand48(0)
    three = multi.align_batch_datums(data, off, None)
    assert np.array_equal(one.rows["hash"], three.rows["hash"]) and np.array_equal(one.rows["rb"], three.rows["rb"])
    assert gpu.session_lrand48() == synth.lrand48_ids(n)[1] if n <= 4096 else True
    t = multi.timing()
    assert t.total > 0 and t.launches > 20
    # the replica answers the row materialisation like the source (its host state travelled with it)
    multi.close()
