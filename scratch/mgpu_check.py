#!/usr/bin/env python
"""Multi-GPU parity check (one process per GPU, launched by torchrun):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scratch/mgpu_check.py [PxQ ...]
For every grid shape given (default: the tester's choice for N ranks and its transpose) runs gemm,
potrf and getrf on seeded matrices distributed 2-D block-cyclically, gathers the result and compares
it on rank 0 with the numpy oracle (identical pivots, factors to 1e-11, tester residuals)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
import slate_b200.host as sl          # noqa: E402
from oracle import slate_oracle as o  # noqa: E402

EPS = np.finfo(np.float64).eps


def gather(M, m, n):
    """Global matrix on every rank: each rank contributes its own tiles, the rest is zero."""
    h = np.zeros((m, n), order="F")
    M.to_host(h)
    t = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()
    dist.all_reduce(t)
    return np.asfortranarray(t.cpu().numpy().T)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    if not shapes:
        p, q = sl.Grid.choose(world)
        shapes = [(p, q)] + ([(q, p)] if p != q else [])
    ok = True
    for (p, q) in shapes:
        grid = sl.Grid.from_torch_distributed(p, q)
        for (n, nb) in [(1024, 128), (1000, 128), (2048, 256)]:
            # ---- getrf
            A = sl.Matrix(n, n, nb, grid).generate("rand", 42)
            piv, info = sl.getrf(A)
            LU = gather(A, n, n)
            if rank == 0:
                A0 = o.generate("rand", n, n, 42)
                LUo, pivo, info_o = o.getrf(A0, nb, 32)
                e = np.abs(LU - LUo).max() / np.abs(LUo).max()
                good = (piv == pivo) and info == info_o == 0 and e <= 1e-11
                print(f"grid {p}x{q} getrf n={n} nb={nb}: pivots {'identical' if piv == pivo else 'DIFFER'}, "
                      f"|LU-LUo|/|LUo|={e:.2e} {'ok' if good else 'FAILED'}", flush=True)
                ok &= good
            # ---- potrf
            H = sl.HermitianMatrix(n, nb, grid).generate("rand_dominant", 42)
            info = sl.potrf(H)
            L = np.tril(gather(H, n, n))
            if rank == 0:
                G = o.generate("rand_dominant", n, n, 42)
                Af = np.tril(G) + np.tril(G, -1).T
                Lo, _ = o.potrf(Af, nb)
                e = np.abs(L - Lo).max() / np.abs(Lo).max()
                good = info == 0 and e <= 64 * EPS
                print(f"grid {p}x{q} potrf n={n} nb={nb}: |L-Lo|/|Lo|={e:.2e} {'ok' if good else 'FAILED'}", flush=True)
                ok &= good
            # ---- gemm
            Am = sl.Matrix(n, n, nb, grid).generate("rand", 1); Bm = sl.Matrix(n, n, nb, grid).generate("rand", 2)
            Cm = sl.Matrix(n, n, nb, grid).generate("rand", 3)
            sl.gemm(3.1, Am, Bm, 2.7, Cm)
            C = gather(Cm, n, n)
            if rank == 0:
                a, b, c = (o.generate("rand", n, n, s) for s in (1, 2, 3))
                ref = o.gemm(3.1, a, b, 2.7, c, nb)
                e = np.abs(C - ref).max() / np.abs(ref).max()
                good = e <= 64 * EPS and o.gemm_check(3.1, a, b, 2.7, c, C) <= 3 * EPS
                print(f"grid {p}x{q} gemm n={n} nb={nb}: err={e:.2e} {'ok' if good else 'FAILED'}", flush=True)
                ok &= good
            dist.barrier()
        grid.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU CHECK", "PASSED" if int(t[0]) else "FAILED", flush=True)
    sys.exit(0 if int(t[0]) else 1)


if __name__ == "__main__":
    main()
