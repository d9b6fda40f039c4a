"""Multi-GPU plumbing: one process per GPU (torch.distributed), K tiled by rows.

The path shards by output rows after a tiny replicated prologue:
every rank packs / relabels the same CSR block and computes rows
[row_block(rank)) of K with `gk_gram(row_begin, row_end)`.  No data-path
collective is needed to *produce* K; `all_gather_rows` assembles the full matrix on
every rank only when a caller asks for it (NCCL over NVLink on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

import numpy as np


TILE = 256  # row blocks of the tiled Gram are multiples of the CTA-pair tile (csrc/comm.h DIST_ALIGN)


def rows_per_rank(n_rows: int, world: int, align: int = 1):
    per = (n_rows + world - 1) // world
    return (per + align - 1) // align * align


def row_block(n_rows: int, rank: int, world: int, align: int = 1):
    """Contiguous block of rows owned by `rank` (ceil split, optionally rounded up to `align` rows per rank;
    trailing ranks may own fewer rows or none).  align=TILE is gk_comm_rows' partition."""
    per = rows_per_rank(n_rows, world, align)
    return min(n_rows, rank * per), min(n_rows, (rank + 1) * per)


def dist_tiles(n_rows: int, rank: int, world: int):
    """The {row, column} tiles gk_gram(GK_DIST) makes `rank` compute (host-only query of the C library)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load_library()
    n = C.c_int64()
    if lib.gk_selftest_dist_tiles(n_rows, world, rank, None, 0, C.byref(n)) != 0:
        raise ValueError(lib.gk_last_error().decode())
    out = np.empty((n.value, 2), dtype=np.int32)
    lib.gk_selftest_dist_tiles(n_rows, world, rank, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
    return out


def comm_init(engine, group=None):
    """Collective: create the engine's communicator (gk_comm_init) with the NCCL id broadcast over torch.distributed."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(world, rank, box[0])
    return rank, world


def all_gather_rows(k_local, n_rows: int, group=None):
    """All-gather row blocks (torch tensor [rows_r, n_cols], same dtype/device on all
    ranks) into the full [n_rows, n_cols] matrix on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = (n_rows + world - 1) // world
    n_cols = k_local.shape[1]
    if k_local.shape[0] == per and k_local.is_contiguous():
        pad = k_local  # already block-sized (rows past this rank's range are sliced off below): no copy
    else:
        pad = torch.zeros((per, n_cols), dtype=k_local.dtype, device=k_local.device)
        pad[: k_local.shape[0]] = k_local
    out = torch.empty((world * per, n_cols), dtype=k_local.dtype, device=k_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_rows]


def gram_rows(compute_rows, n_rows: int, rank: int, world: int):
    """Run `compute_rows(row_begin, row_end) -> ndarray [rows, n]` for this rank's block."""
    rb, re_ = row_block(n_rows, rank, world)
    k = compute_rows(rb, re_)
    assert k.shape[0] == re_ - rb
    return rb, re_, k


def wl_gram_rows(engine, graph_ptr, row_ptr, col_idx, labels, n_iter, rank, world, dtype=np.float64, out=None,
                 normalize=False):
    """This rank's row block of the WL-subtree Gram matrix of the packed graphs."""
    n = len(graph_ptr) - 1
    rb, re_ = row_block(n, rank, world)
    engine.pack(graph_ptr, row_ptr, col_idx, labels)
    st = engine.wl_features(n_iter)
    k, xd, _ = engine.gram(n, normalize=normalize, nan_to_num=normalize, dtype=dtype,
                           row_range=(rb, re_) if world > 1 else None, out=out, stats=st)
    return rb, re_, k, xd, st
