"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: row partition + all-gather of
row blocks reproduce the single-process matrix.  The per-rank compute is the numpy model
of the device pipeline (tests/blockref.py) restricted to the rank's rows."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import blockref
        from grakel_b200.dist import all_gather_rows, gram_rows, row_block
        from grakel_b200.packing import label_ids, pack
        from oracle.gk_oracle import gen

        X = gen(37, 12, 4)  # 37 rows over 2 ranks: uneven blocks
        b = pack(X, "wl", len_ok=lambda n: n >= 2)
        ids, _ = label_ids(b.labels, None)
        K_full, _, _ = blockref.wl_gram_block(b, ids, 3)
        rb, re_, k = gram_rows(lambda r0, r1: K_full[r0:r1], b.n_graphs, rank, world)
        assert (rb, re_) == row_block(37, rank, world)
        K = all_gather_rows(torch.from_numpy(np.ascontiguousarray(k)), b.n_graphs)
        ok = bool(np.array_equal(K.numpy(), K_full))
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == world
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_row_blocks_cover_everything():
    from grakel_b200.dist import row_block
    for n in (1, 2, 7, 10000, 10001):
        for w in (1, 2, 4, 8):
            blocks = [row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))


@pytest.mark.parametrize("n,world", [(1, 2), (255, 2), (256, 2), (1000, 3), (3000, 4), (10000, 8), (14142, 2), (28284, 8)])
def test_dist_tiles_cover_every_block_exactly_once(n, world):
    """gk_gram(GK_DIST): every rank computes the upper triangle of its diagonal block plus its parity share of the
    rectangles it shares with the other ranks; the direct block of a tile lands in the computing rank's rows, the
    mirrored block in the owner of the tile's columns.  Together they must write every 256 x 256 block of K
    exactly once (diagonal tiles: once), and the load must be balanced."""
    from grakel_b200.dist import TILE, dist_tiles, row_block, rows_per_rank
    per = rows_per_rank(n, world, TILE)
    nt = (n + TILE - 1) // TILE
    written = np.zeros((nt, nt), dtype=np.int32)
    loads = []
    for r in range(world):
        rb, re_ = row_block(n, r, world, TILE)
        t = dist_tiles(n, r, world)
        loads.append(len(t))
        for x, y in t.tolist():
            assert x % TILE == 0 and y % TILE == 0 and rb <= x < re_ and y < n
            ti, tj = x // TILE, y // TILE
            written[ti, tj] += 1
            if ti != tj:
                written[tj, ti] += 1  # the mirrored half, stored into the owner of rows [y, y + 256)
                owner = y // per
                orb, ore = row_block(n, owner, world, TILE)
                assert orb <= y < ore
    assert np.all(written == 1), "a block of K is written twice or never"
    busy = [l for l in loads if l]
    if n >= 3000:
        assert max(busy) <= 1.35 * (sum(loads) / world) + 2


def test_gloo_world2_row_tiling_and_gather():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
