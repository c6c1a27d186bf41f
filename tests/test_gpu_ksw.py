"""the warp-cooperative ksw_extend2 / ksw_global2 kernels vs the scalar oracle
on random job lists (bit-exact: score, qle, tle, gtle, gscore, max_off; score and CIGAR)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from bioseqdb_b200 import _lib
from helpers import to_bsq

pytestmark = pytest.mark.gpu


def _jobs(rng, n, qmax, err, n_frac=0.0):
    qs, ts = [], []
    for _ in range(n):
        ql = int(rng.integers(1, qmax + 1))
        q = rng.integers(0, 4, size=ql).astype(np.uint8)
        t = []
        for b in q:  # target = query with substitutions / indels
            r = rng.random()
            if r < err / 3:
                continue
            if r < 2 * err / 3:
                t.append(int(rng.integers(0, 4)))
            t.append(int((b + int(rng.integers(1, 4))) % 4) if r > 1 - err / 3 else int(b))
        t = np.array(t + list(rng.integers(0, 4, size=int(rng.integers(0, 30)))), dtype=np.uint8)
        if rng.random() < 0.1:
            t = t[:int(rng.integers(0, len(t) + 1))]
        if n_frac:
            q[rng.random(ql) < n_frac] = 4
        qs.append(q)
        ts.append(t)
    q_off = np.zeros(n + 1, dtype=np.uint64); q_off[1:] = np.cumsum([len(x) for x in qs])
    t_off = np.zeros(n + 1, dtype=np.uint64); t_off[1:] = np.cumsum([len(x) for x in ts])
    return qs, ts, np.concatenate(qs), np.concatenate(ts + [np.zeros(1, np.uint8)]), q_off, t_off


@pytest.mark.parametrize("opts_fn,qmax,err", [(O.sql_default_opts, 131, 0.03), (O.canonical_opts, 131, 0.03), (O.canonical_opts, 600, 0.12),
                                              (O.sql_default_opts, 40, 0.3)])
def test_ksw_extend_parity(gpu_lib, opts_fn, qmax, err):
    rng = np.random.default_rng(qmax * 7 + int(err * 100))
    opts = opts_fn(1)
    n = 1500
    qs, ts, qcat, tcat, q_off, t_off = _jobs(rng, n, qmax, err, n_frac=0.01)
    w = rng.choice([100, 200, 5, 30], size=n).astype(np.int32)
    eb = rng.choice([5, 0], size=n).astype(np.int32)
    h0 = rng.integers(1, 150, size=n).astype(np.int32)
    out = np.zeros((n, 6), dtype=np.int32)
    b = to_bsq(opts)
    _lib.check(gpu_lib.bsq_debug_ksw_extend(C.byref(b), 0, n, _lib.ptr(qcat), _lib.ptr(q_off), _lib.ptr(tcat), _lib.ptr(t_off),
                                            _lib.ptr(w), _lib.ptr(eb), _lib.ptr(h0), _lib.ptr(out)))
    L = O.lib()
    for i in range(n):
        o5 = np.zeros(5, dtype=np.int32)
        q = np.ascontiguousarray(qs[i]); t = np.ascontiguousarray(ts[i]) if len(ts[i]) else np.zeros(1, np.uint8)
        sc = L.orc_ksw_extend2(len(qs[i]), O._ptr(q), len(ts[i]), O._ptr(t), C.byref(opts), int(w[i]), int(eb[i]), int(h0[i]), O._ptr(o5))
        assert [sc] + o5.tolist() == out[i].tolist(), (i, len(qs[i]), len(ts[i]), int(w[i]), int(h0[i]))


@pytest.mark.parametrize("opts_fn,qmax,err,reversed_q", [(O.sql_default_opts, 136, 0.03, 0), (O.canonical_opts, 136, 0.03, 1), (O.canonical_opts, 136, 0.12, 0),
                                                         (O.sql_default_opts, 40, 0.3, 1), (O.sql_default_opts, 9, 0.1, 1)])
def test_ksw_extend_thread_parity(gpu_lib, opts_fn, qmax, err, reversed_q):
    """The thread-per-extension kernel of the production pre-pass (ksw_thread.cuh) against the scalar oracle, through both of
    its query loaders (forward = right extension, back to front = left extension)."""
    rng = np.random.default_rng(qmax * 11 + int(err * 100) + reversed_q)
    opts = opts_fn(1)
    n = 3000
    qs, ts, qcat, tcat, q_off, t_off = _jobs(rng, n, qmax, err, n_frac=0.02)
    w = rng.choice([100, 200, 5, 30], size=n).astype(np.int32)
    eb = rng.choice([5, 0], size=n).astype(np.int32)
    h0 = rng.integers(1, 150, size=n).astype(np.int32)
    out = np.zeros((n, 6), dtype=np.int32)
    b = to_bsq(opts)
    qdev = np.concatenate([x[::-1] for x in qs]) if reversed_q else qcat
    qdev = np.ascontiguousarray(qdev)
    _lib.check(gpu_lib.bsq_debug_ksw_extend_thread(C.byref(b), 0, n, _lib.ptr(qdev), _lib.ptr(q_off), _lib.ptr(tcat), _lib.ptr(t_off),
                                                   _lib.ptr(w), _lib.ptr(eb), _lib.ptr(h0), _lib.ptr(out), reversed_q))
    L = O.lib()
    for i in range(n):
        o5 = np.zeros(5, dtype=np.int32)
        q = np.ascontiguousarray(qs[i]); t = np.ascontiguousarray(ts[i]) if len(ts[i]) else np.zeros(1, np.uint8)
        sc = L.orc_ksw_extend2(len(qs[i]), O._ptr(q), len(ts[i]), O._ptr(t), C.byref(opts), int(w[i]), int(eb[i]), int(h0[i]), O._ptr(o5))
        assert [sc] + o5.tolist() == out[i].tolist(), (i, len(qs[i]), len(ts[i]), int(w[i]), int(h0[i]))


@pytest.mark.parametrize("opts_fn,qmax,err", [(O.sql_default_opts, 150, 0.03), (O.canonical_opts, 150, 0.05), (O.canonical_opts, 500, 0.1)])
def test_ksw_global_parity(gpu_lib, opts_fn, qmax, err):
    rng = np.random.default_rng(qmax + int(err * 1000))
    opts = opts_fn(1)
    n = 800
    qs, ts, qcat, tcat, q_off, t_off = _jobs(rng, n, qmax, err, n_frac=0.01)
    keep = [i for i in range(n) if len(ts[i]) > 0]
    w = np.array([abs(len(ts[i]) - len(qs[i])) + int(rng.integers(3, 40)) for i in range(n)], dtype=np.int32)
    cap = 2 * qmax + 64
    sc = np.zeros(n, dtype=np.int32); cig = np.zeros((n, cap), dtype=np.uint32); nc = np.zeros(n, dtype=np.int32)
    b = to_bsq(opts)
    _lib.check(gpu_lib.bsq_debug_ksw_global(C.byref(b), 0, n, _lib.ptr(qcat), _lib.ptr(q_off), _lib.ptr(tcat), _lib.ptr(t_off),
                                            _lib.ptr(w), _lib.ptr(sc), _lib.ptr(cig), cap, _lib.ptr(nc)))
    L = O.lib()
    for i in keep:
        oc = np.zeros(cap, dtype=np.uint32); on = C.c_int()
        q = np.ascontiguousarray(qs[i]); t = np.ascontiguousarray(ts[i])
        osc = L.orc_ksw_global2(len(q), O._ptr(q), len(t), O._ptr(t), C.byref(opts), int(w[i]), O._ptr(oc), cap, C.byref(on))
        assert osc == sc[i], (i, osc, sc[i])
        assert on.value == nc[i] and np.array_equal(oc[:on.value], cig[i, :nc[i]]), (i, O.cigar_str(oc[:on.value]), O.cigar_str(cig[i, :nc[i]]))
