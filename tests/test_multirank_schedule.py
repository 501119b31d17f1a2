"""CPU test (world_size 2, gloo) of the N > 1 SCHEDULES the C++ runtime runs over NCCL, restated with numpy tiles and
torch.distributed collectives on the same communicators (world, process-row group, process-column group), tile owners
and broadcast roots taken from the library's own host-only tile map (sb200_tile_rank):

  * potrf_driver (csrc/runtime.cu):  diagonal tile on its owner -> L_kk down the process column -> panel solve on the
    owners -> every rank receives the whole factored block column (p broadcasts over world, root = owner of the first
    tile of each process row) -> local trailing update;
  * sweep_dist (csrc/solve_dist.cu): replicated right-hand sides, left-looking partial sums on the ranks that own block
    row (NoTrans) / block column (ConjTrans) i, all-reduce inside that process row / column, block solve on the owner
    of T(i,i), world broadcast of the solved block.

The result must equal the serial oracle (oracle/slate_oracle.py: potrf, potrs) on 1x2 and 2x1 grids.  This pins the
communication pattern and ownership arithmetic of the multi-rank drivers without a GPU; the CUDA path itself is
checked on GPUs by scratch/mgpu_check.py."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, p, q, n, nb, nrhs, out_dir):
    import torch
    import torch.distributed as dist
    from scipy.linalg import solve_triangular
    import slate_b200.host as sl
    from oracle import slate_oracle as o
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    prow, pcol = rank % p, rank // p
    owner = lambda i, j: sl.tile_rank(p, q, i, j)
    row_groups = [dist.new_group([r + c * p for c in range(q)]) for r in range(p)]       # same prow
    col_groups = [dist.new_group([r + c * p for r in range(p)]) for c in range(q)]       # same pcol
    nt = -(-n // nb)
    sz = lambda i: min(nb, n - i * nb)

    def bcast(arr, src, group=None):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        dist.broadcast(t, src, group=group)
        return t.numpy()

    G = o.generate("rand_dominant", n, n, 42)
    Afull = o.he_full(G)
    A = {(i, j): Afull[i * nb:i * nb + sz(i), j * nb:j * nb + sz(j)].copy()
         for j in range(nt) for i in range(j, nt) if owner(i, j) == rank}

    # ---------------- potrf_driver schedule
    for k in range(nt):
        own = owner(k, k)
        in_col = pcol == k % q
        Lkk = np.zeros((sz(k), sz(k)))
        if rank == own:
            A[(k, k)] = np.linalg.cholesky(np.tril(A[(k, k)]) + np.tril(A[(k, k)], -1).T)
            Lkk = A[(k, k)]
        if in_col and p > 1:
            Lkk = bcast(Lkk, own, col_groups[pcol])                        # L_kk down the process column
        if in_col:
            for i in range(k + 1, nt):
                if owner(i, k) == rank:
                    A[(i, k)] = solve_triangular(Lkk, A[(i, k)].T, lower=True).T
        # every rank receives the factored block column: one broadcast per process row, root = owner of its first tile
        panel = {}
        for r in range(p):
            i0 = k + 1 + ((r - (k + 1)) % p + p) % p
            if i0 >= nt:
                continue
            rows = list(range(i0, nt, p))
            root = owner(i0, k)
            stack = np.concatenate([A[(i, k)] if rank == root else np.zeros((sz(i), sz(k))) for i in rows], axis=0)
            stack = bcast(stack, root)
            o0 = 0
            for i in rows:
                panel[i] = stack[o0:o0 + sz(i)]; o0 += sz(i)
        for j in range(k + 1, nt):
            for i in range(j, nt):
                if owner(i, j) == rank:
                    A[(i, j)] -= panel[i] @ panel[j].T
    Lo, info = o.potrf(Afull, nb)
    assert info == 0
    for (i, j), t in A.items():
        ref = Lo[i * nb:i * nb + sz(i), j * nb:j * nb + sz(j)]
        tt = np.tril(t) if i == j else t
        assert np.abs(tt - ref).max() <= 64 * np.finfo(float).eps * np.abs(Lo).max(), (i, j)

    # ---------------- sweep_dist schedule (potrs): replicated X, lower NoTrans forward then lower Trans backward
    X = o.generate("rand", n, nrhs, 43)                                    # replicated on every rank
    blk = lambda i: slice(i * nb, i * nb + sz(i))
    for trans in (False, True):
        order = range(nt) if not trans else range(nt - 1, -1, -1)
        for i in order:
            ks = range(0, i) if not trans else range(i + 1, nt)
            in_set = (prow == i % p) if not trans else (pcol == i % q)
            part = np.zeros((sz(i), nrhs))
            if in_set and len(ks) > 0:
                for k in ks:
                    key = (i, k) if not trans else (k, i)
                    if owner(*key) == rank:
                        part += (A[key] if not trans else A[key].T) @ X[blk(k)]
                grp = row_groups[prow] if not trans else col_groups[pcol]
                t = torch.from_numpy(part); dist.all_reduce(t, group=grp); part = t.numpy()
            own = owner(i, i)
            if rank == own:
                T = np.tril(A[(i, i)])
                X[blk(i)] = solve_triangular(T, X[blk(i)] - part, lower=True, trans=1 if trans else 0)
            X[blk(i)] = bcast(X[blk(i)], own)
    Xo = o.potrs(Lo, o.generate("rand", n, nrhs, 43), nb)
    assert np.abs(X - Xo).max() <= 200 * np.finfo(float).eps * np.abs(Xo).max()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("p,q,n,nb", [(1, 2, 640, 128), (2, 1, 600, 128)])
def test_potrf_and_replicated_rhs_sweeps_world2_gloo(tmp_path, p, q, n, nb):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, p, q, n, nb, 5, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
