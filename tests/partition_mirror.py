"""TEST SCAFFOLDING (not product code): a numpy restatement of the row partition and ParCSR split.

The product performs these steps on the device (csrc/hdk_csr.cu: k_split_count / k_split_fill,
csrc/hdk_comm.cu: build_halo_plan; the device code itself is exercised multi-rank by
tests/test_gpu_multi.py, also on a 1-GPU box).  This mirror only lets the world_size-2 rendezvous
and halo bookkeeping be exercised on CPU with torch.distributed's gloo backend
(tests/test_multirank_gloo.py).  Partition contract: the
reference's, include/HYPREDRV.h:836-839 -- contiguous inclusive row ranges in rank order.
"""
from __future__ import annotations

import numpy as np


def slab_range(n_global: int, rank: int, world: int):
    """Inclusive [row_start, row_end] of `rank` for near-equal contiguous slabs."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0) - 1


def row_starts(n_global: int, world: int):
    return [slab_range(n_global, r, world)[0] for r in range(world)] + [n_global]


def split_diag_offd(indptr, cols, data, row_start: int, row_end: int):
    """Global-column CSR rows -> (diag CSR with local columns, diagonal swapped first),
    (offd CSR with compressed columns), col_map_offd (sorted global ids)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    data = np.asarray(data, dtype=np.float64)
    n = indptr.size - 1
    local = (cols >= row_start) & (cols <= row_end)
    d_ptr, o_ptr = np.zeros(n + 1, dtype=np.int64), np.zeros(n + 1, dtype=np.int64)
    d_col, d_val, o_gcol, o_val = [], [], [], []
    for r in range(n):
        s, e = indptr[r], indptr[r + 1]
        m = local[s:e]
        dc, dv = (cols[s:e][m] - row_start).copy(), data[s:e][m].copy()
        hit = np.nonzero(dc == r)[0]
        if hit.size and hit[0] != 0:                    # hypre's swap (not a rotation)
            k = hit[0]
            dc[[0, k]] = dc[[k, 0]]
            dv[[0, k]] = dv[[k, 0]]
        d_col.append(dc); d_val.append(dv)
        o_gcol.append(cols[s:e][~m]); o_val.append(data[s:e][~m])
        d_ptr[r + 1] = d_ptr[r] + dc.size
        o_ptr[r + 1] = o_ptr[r] + (e - s - dc.size)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)
    o_g = cat(o_gcol, np.int64)
    col_map = np.unique(o_g)
    diag = (d_ptr, cat(d_col, np.int64), cat(d_val, np.float64))
    offd = (o_ptr, np.searchsorted(col_map, o_g).astype(np.int64), cat(o_val, np.float64))
    return diag, offd, col_map


def halo_plan(col_map, starts, rank: int):
    """Who owns my halo columns: {owner: (offset, count)} over the sorted col_map."""
    col_map = np.asarray(col_map, dtype=np.int64)
    owners = np.searchsorted(np.asarray(starts, dtype=np.int64), col_map, side="right") - 1
    plan = {}
    for r in np.unique(owners):
        idx = np.nonzero(owners == r)[0]
        if int(r) == rank:
            raise ValueError("a halo column is owned by the requesting rank")
        plan[int(r)] = (int(idx[0]), int(idx.size))
    return plan
