"""Multi-GPU plumbing: one process per GPU, the FM-index is built once and broadcast
(NCCL over NVLink on the GPU box, gloo in the CPU tests), reads are split in contiguous blocks so that read
i keeps lrand48 id i, results are gathered in read order. There is no collective in the per-read path."""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of reads for `rank` (same rule as the oracle's thread sharding)."""
    return n * rank // world, n * (rank + 1) // world


def shard_reads(seqs: np.ndarray, offs: np.ndarray, ids: np.ndarray, rank: int, world: int):
    n = len(offs) - 1
    lo, hi = shard_bounds(n, rank, world)
    o = offs[lo:hi + 1]
    return seqs[int(o[0]):int(o[-1])], (o - o[0]).astype(np.uint64), ids[lo:hi], (lo, hi)


class _CudaArray:
    """Raw device pointer exposed through __cuda_array_interface__ so that torch can wrap it."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_meta(meta_bytes: bytes | None, nbytes: int, dist, device, src: int = 0) -> bytes:
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if meta_bytes is not None:
        t.copy_(torch.frombuffer(bytearray(meta_bytes), dtype=torch.uint8))
    dist.broadcast(t, src)
    return t.cpu().numpy().tobytes()


def broadcast_index(ix, rank: int, dist, src: int = 0) -> int:
    """Rank `src` holds a built BwaIndex; the others allocate replicas and receive every index array.
    Returns the number of bytes broadcast."""
    import torch
    from . import _lib
    from ._lib import BsqMeta
    raw = broadcast_meta(bytes(ix.meta()) if rank == src else None, C.sizeof(BsqMeta), dist, "cuda", src)
    m = BsqMeta.from_buffer_copy(raw)
    if rank != src:
        _lib.check(ix.L.bsq_index_alloc_replica(ix.h, C.byref(m)))
    total = 0
    for what in range(_lib.ARR_COUNT):
        n = int(m.arr_bytes[what])
        if n == 0:
            continue
        p = C.c_void_p()
        _lib.check(ix.L.bsq_index_device_ptr(ix.h, what, C.byref(p)))
        t = torch.as_tensor(_CudaArray(p.value, n), device="cuda")
        dist.broadcast(t, src)
        total += n
    torch.cuda.synchronize()
    # host-only state (the reference rows' ambiguity holes): every replica must overlay the same holes on ref_subseq
    st = ix.host_state() if rank == src else None
    nb = torch.tensor([len(st) if st is not None else 0], dtype=torch.int64, device="cuda")
    dist.broadcast(nb, src)
    st = broadcast_meta(st, int(nb.item()), dist, "cuda", src)
    if rank != src:
        ix.replica_finish(st)
    return total + int(nb.item())


def broadcast_host_arrays(arrays: dict | None, names, dist, src: int = 0) -> dict:
    """CPU (gloo) version used by the tests: broadcast a dict of numpy arrays from `src`."""
    import torch
    out = {}
    for name in names:
        if arrays is not None:
            a = np.ascontiguousarray(arrays[name])
            hdr = torch.tensor([a.nbytes, a.dtype.itemsize], dtype=torch.int64)
        else:
            hdr = torch.zeros(2, dtype=torch.int64)
        dist.broadcast(hdr, src)
        nbytes, itemsize = int(hdr[0]), int(hdr[1])
        t = torch.zeros(nbytes, dtype=torch.uint8)
        if arrays is not None:
            t.copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
        dist.broadcast(t, src)
        out[name] = t.numpy().copy().view({1: np.uint8, 4: np.uint32, 8: np.uint64}[itemsize])
    return out


def gather_rows(row_off: np.ndarray, rows: np.ndarray, cigar: np.ndarray, dist, rank: int, world: int, dst: int = 0):
    """Concatenate per-rank results in read order on `dst` (host side, after the download)."""
    import torch
    payload = [row_off, rows, cigar]
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    offs = [np.zeros(1, dtype=np.uint64)]
    all_rows, all_cig = [], []
    base_rows = 0
    base_cig = 0
    for ro, rw, cg in gathered:
        offs.append(ro[1:].astype(np.uint64) + np.uint64(base_rows))
        rw = rw.copy()
        rw["cigar_off"] += np.uint32(base_cig)
        all_rows.append(rw)
        all_cig.append(cg)
        base_rows += len(rw)
        base_cig += len(cg)
    return np.concatenate(offs), np.concatenate(all_rows), np.concatenate(all_cig)


def gather_rows_host(row_off: np.ndarray, rows: np.ndarray, cigar: np.ndarray, dist, group, rank: int, world: int, dst: int = 0):
    """gather_rows over a CPU (gloo) group with plain byte tensors instead of pickled objects: the results are already in host
    memory (one process per GPU), so this is a host-to-host copy of every rank's row offsets, rows and CIGAR words to `dst`."""
    import torch
    sizes = torch.tensor([row_off.nbytes, rows.nbytes, cigar.nbytes], dtype=torch.int64)
    all_sizes = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    parts = []
    for k, arr in enumerate((row_off, rows, cigar)):
        mine = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()) if arr.nbytes else torch.zeros(0, dtype=torch.uint8)
        cap = int(max(int(s[k]) for s in all_sizes))
        buf = torch.zeros(cap, dtype=torch.uint8)
        buf[:mine.numel()] = mine
        got = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)] if rank == dst else None
        dist.gather(buf, got, dst=dst, group=group)
        if rank == dst:
            parts.append([g[:int(all_sizes[r][k])].numpy() for r, g in enumerate(got)])
    if rank != dst:
        return None
    offs = [np.zeros(1, dtype=np.uint64)]
    all_rows, all_cig = [], []
    base_rows = base_cig = 0
    for r in range(world):
        ro = parts[0][r].view(np.uint64)
        rw = parts[1][r].view(rows.dtype).copy()
        cg = parts[2][r].view(np.uint32)
        offs.append(ro[1:] + np.uint64(base_rows))
        rw["cigar_off"] += np.uint32(base_cig)
        all_rows.append(rw)
        all_cig.append(cg)
        base_rows += len(rw)
        base_cig += len(cg)
    return np.concatenate(offs), np.concatenate(all_rows), np.concatenate(all_cig)
