"""CPU tests (world_size 2, gloo) of the host-side logic of the N > 1 path: the 2-D block-cyclic tile map the C++
runtime uses (tileRank = (i % p) + (j % q) * p, include/slate/func.hh:96-104), the per-rank packed local tile
order of Matrix.from_host_local / to_host_local, and the grid choice bench.py makes.  No GPU, no compute calls:
the entries exercised here are the host-only functions of the C ABI."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, p, q, kind, m, n, nb, out_dir):
    import torch
    import torch.distributed as dist
    import slate_b200.host as sl
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mt, nt = -(-m // nb), -(-n // nb)
    K = ord(kind)
    mine = [(i, j) for j in range(nt) for i in range(mt)
            if (kind == "G" or i >= j) and sl.tile_rank(p, q, i, j) == rank]
    cnt = sl.local_tile_count(K, p, q, rank, m, n, nb)
    assert cnt == len(mine), (cnt, len(mine))
    # packed local order: slots 0..cnt-1, each used once, ordered by (local block column, local block row)
    slots = [sl.local_tile_index(K, p, q, m, n, nb, i, j) for (i, j) in mine]
    assert sorted(slots) == list(range(cnt))
    order = [t for _, t in sorted(zip(slots, mine))]
    assert order == sorted(mine, key=lambda t: (t[1], t[0]))
    # every stored tile is owned by exactly one rank: sum of counts over ranks == number of stored tiles
    tot = torch.tensor([cnt], dtype=torch.int64)
    dist.all_reduce(tot)
    stored = sum(1 for j in range(nt) for i in range(mt) if kind == "G" or i >= j)
    assert int(tot) == stored
    # owners agree across ranks (each rank publishes the owner map it computed)
    owner = torch.tensor([sl.tile_rank(p, q, i, j) for j in range(nt) for i in range(mt)], dtype=torch.int64)
    ref = owner.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(owner, ref)
    # panel broadcast roots of the potrf / herk drivers: tiles (i, k), i % p == r, are contiguous slots on their root
    for k in range(min(nt, 3)):
        for r in range(p):
            rows = [i for i in range(k + 1, mt) if i % p == r and (kind == "G" or i >= k)]
            if not rows or sl.tile_rank(p, q, rows[0], k) != rank:
                continue
            s = [sl.local_tile_index(K, p, q, m, n, nb, i, k) for i in rows]
            assert s == list(range(s[0], s[0] + len(s)))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("p,q,kind,m,n,nb", [(1, 2, "G", 1000, 1300, 128), (2, 1, "H", 1024, 1024, 128),
                                             (1, 2, "H", 900, 900, 256), (2, 1, "G", 640, 384, 128)])
def test_tile_map_partition_world2_gloo(tmp_path, p, q, kind, m, n, nb):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, p, q, kind, m, n, nb, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_grid_choice_matches_reference_tester():
    """test/test.cc:738-747: p x q as square as possible with p <= q."""
    import slate_b200.host as sl
    assert [sl.Grid.choose(w) for w in (1, 2, 4, 8, 6, 16)] == [(1, 1), (1, 2), (2, 2), (2, 4), (2, 3), (4, 4)]


def test_tile_map_matches_numpy_restatement():
    import slate_b200.host as sl
    for p, q in ((2, 4), (1, 1), (3, 2)):
        for kind in "GH":
            m = n = 1500; nb = 128
            mt = -(-m // nb)
            tot = sum(sl.local_tile_count(ord(kind), p, q, r, m, n, nb) for r in range(p * q))
            assert tot == (mt * mt if kind == "G" else mt * (mt + 1) // 2)
            for i in range(mt):
                for j in range(mt):
                    assert sl.tile_rank(p, q, i, j) == (i % p) + (j % q) * p
    assert sl.local_tile_index(ord("H"), 2, 2, 512, 512, 128, 0, 1) == -1        # upper tile: not stored
