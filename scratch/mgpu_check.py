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
    h = np.zeros((m, n), dtype=M.dtype, order="F")
    M.to_host(h)
    if np.iscomplexobj(h):
        parts = []
        for comp in (h.real, h.imag):
            t = torch.from_numpy(np.ascontiguousarray(comp.T)).cuda()
            dist.all_reduce(t)
            parts.append(t.cpu().numpy().T)
        return np.asfortranarray(parts[0] + 1j * parts[1])
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
        # ---- sb200_bcast_tiles (the listBcast hook): three ranges from different roots, one of them large enough for
        #      the scatter + all-gather form, one with an odd byte count
        if os.environ.get("MGPU_BCAST", "1") != "0":
            good = True
            rngs, bufs = [], []
            for t, (nbytes, root) in enumerate([(8 << 20, 0), (777, world - 1), ((5 << 20) + 24, world // 2)]):
                srcb = torch.full((nbytes,), 17 + t, dtype=torch.uint8, device="cuda") if rank == root else None
                dstb = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
                if rank == root:
                    srcb[::3] = (torch.arange(0, (nbytes + 2) // 3, device="cuda") % 251).to(torch.uint8)
                bufs.append((srcb, dstb, nbytes, root, t))
                rngs.append((srcb.data_ptr() if rank == root else 0, dstb.data_ptr(), nbytes, root))
            grid.bcast_tiles(rngs)
            torch.cuda.synchronize()
            for (srcb, dstb, nbytes, root, t) in bufs:
                ref = torch.full((nbytes,), 17 + t, dtype=torch.uint8, device="cuda")
                ref[::3] = (torch.arange(0, (nbytes + 2) // 3, device="cuda") % 251).to(torch.uint8)
                good &= bool(torch.equal(dstb, ref))
            flag = torch.tensor([1 if good else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"grid {p}x{q} bcast_tiles: {'ok' if int(flag[0]) else 'FAILED'}", flush=True)
            ok &= bool(int(flag[0]))
        sizes = [(1024, 128), (1000, 128), (2048, 256)]
        if os.environ.get("MGPU_SIZES"):
            sizes = [tuple(int(x) for x in a.split("x")) for a in os.environ["MGPU_SIZES"].split(",")]
        for (n, nb) in sizes:
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
            # ---- round-1 widening on the grid: complex herk / potrf / gemm, FP32 potrf on the tcgen05 kernel
            if n != 2048:
                k = n // 2
                Az = sl.Matrix(n, k, nb, grid, "z").generate("rand", 5)
                Cz = sl.HermitianMatrix(n, nb, grid, dtype="z").generate("rand", 6)
                sl.herk(-1.0, Az, 2.0, Cz)
                out = np.tril(gather(Cz, n, n))
                Hz = sl.HermitianMatrix(n, nb, grid, dtype="z").generate("rand_dominant", 42)
                iz = sl.potrf(Hz)
                Lz = np.tril(gather(Hz, n, n))
                Bz = sl.Matrix(k, n, nb, grid, "z").generate("rand", 7)
                Gz = sl.Matrix(n, n, nb, grid, "z").generate("rand", 8)
                sl.gemm(3.1 + 1.4j, Az, Bz, 2.7 + 1.7j, Gz)
                gz = gather(Gz, n, n)
                Hs = sl.HermitianMatrix(n, nb, grid, dtype="s").generate("rand_dominant", 42)
                i_s = sl.potrf(Hs, {"tensor_core_fp32": True})
                Ls = np.tril(gather(Hs, n, n)).astype(np.float64)
                if rank == 0:
                    a = o.generate("rand", n, k, 5, np.complex128)
                    ref = np.tril(o.herk(-1.0, a, 2.0, o.generate("rand", n, n, 6, np.complex128), nb))
                    e1 = np.abs(out - ref).max() / np.abs(ref).max()
                    Lo, _ = o.potrf(o.he_full(o.generate("rand_dominant", n, n, 42, np.complex128)), nb)
                    e2 = np.abs(Lz - Lo).max() / np.abs(Lo).max()
                    refg = o.gemm(3.1 + 1.4j, a, o.generate("rand", k, n, 7, np.complex128), 2.7 + 1.7j,
                                  o.generate("rand", n, n, 8, np.complex128), nb)
                    e3 = np.abs(gz - refg).max() / np.abs(refg).max()
                    Lso, _ = o.potrf(o.he_full(o.generate("rand_dominant", n, n, 42, np.float32).astype(np.float64)), nb)
                    e4 = np.abs(Ls - Lso).max() / np.abs(Lso).max()
                    good = (e1 <= 256 * EPS and e2 <= 64 * EPS and iz == 0 and e3 <= 256 * EPS
                            and i_s == 0 and e4 <= 64 * float(np.finfo(np.float32).eps))
                    print(f"grid {p}x{q} n={n} nb={nb}: zherk err={e1:.2e} zpotrf err={e2:.2e} zgemm err={e3:.2e} "
                          f"spotrf(tcgen05) err={e4:.2e} {'ok' if good else 'FAILED'}", flush=True)
                    ok &= good
            # ---- SURVEY 8(f) items 2-3 on the grid (written after round 1's GPU budget was spent): her2k, complex-symmetric
            #      syrk / syr2k, getrf_nopiv.  MGPU_WIDEN=0 skips them.
            if os.environ.get("MGPU_WIDEN", "1") != "0" and n != 2048:
                k = n // 2
                al, be = 3.1 + 1.4j, 2.7 + 1.7j
                Az = sl.Matrix(n, k, nb, grid, "z").generate("rand", 5)
                Bz = sl.Matrix(n, k, nb, grid, "z").generate("rand", 9)
                outs = []
                for name in ("her2k", "syrk", "syr2k"):
                    Cz = sl.HermitianMatrix(n, nb, grid, dtype="z").generate("rand", 6)
                    if name == "her2k":
                        sl.her2k(al, Az, Bz, 2.0, Cz)
                    elif name == "syrk":
                        sl.syrk(al, Az, be, Cz)
                    else:
                        sl.syr2k(al, Az, Bz, be, Cz)
                    outs.append(np.tril(gather(Cz, n, n)))
                Gn = sl.Matrix(n, n, nb, grid).generate("rand_dominant", 42)
                inp = sl.getrf_nopiv(Gn)
                lun = gather(Gn, n, n)
                if rank == 0:
                    a, b = o.generate("rand", n, k, 5, np.complex128), o.generate("rand", n, k, 9, np.complex128)
                    c = np.tril(o.generate("rand", n, n, 6, np.complex128))
                    refs = [np.tril(o.her2k(al, a, b, 2.0, c, nb)), np.tril(o.syrk(al, a, be, c, nb)),
                            np.tril(o.syr2k(al, a, b, be, c, nb))]
                    errs = [np.abs(x - r).max() / np.abs(r).max() for x, r in zip(outs, refs)]
                    LUo, info_o = o.getrf_nopiv(o.generate("rand_dominant", n, n, 42), nb)
                    en = np.abs(lun - LUo).max() / np.abs(LUo).max()
                    good = all(e <= 256 * EPS for e in errs) and en <= 64 * EPS and inp == info_o == 0
                    print(f"grid {p}x{q} n={n} nb={nb}: zher2k err={errs[0]:.2e} zsyrk err={errs[1]:.2e} zsyr2k err={errs[2]:.2e} "
                          f"getrf_nopiv err={en:.2e} {'ok' if good else 'FAILED'}", flush=True)
                    ok &= good
            # ---- LU with tournament pivoting on the grid (written after round 2's GPU budget was spent: the tournament
            #      itself is validated on one rank through SB200_TNT_RANKS; this checks it behind the real gather /
            #      broadcast).  Participants per panel = the grid's process rows.  MGPU_TNT=0 skips it.
            if os.environ.get("MGPU_TNT", "1") != "0":
                At = sl.Matrix(n, n, nb, grid).generate("rand", 42)
                pivt, infot = sl.getrf_tntpiv(At)
                LUt = gather(At, n, n)
                if rank == 0:
                    LUo, pivo, info_o = o.getrf_tntpiv(o.generate("rand", n, n, 42), nb, 32, ranks=p)
                    e = np.abs(LUt - LUo).max() / np.abs(LUo).max()
                    good = (pivt == pivo) and infot == info_o == 0 and e <= 1e-11
                    print(f"grid {p}x{q} getrf_tntpiv n={n} nb={nb}: pivots {'identical' if pivt == pivo else 'DIFFER'}, "
                          f"|LU-LUo|/|LUo|={e:.2e} {'ok' if good else 'FAILED'}", flush=True)
                    ok &= good
            # ---- solve path on the grid (replicated right-hand sides, solve_dist.cu): potrs and posv_mixed
            if os.environ.get("MGPU_SOLVE", "1") != "0":
                Hd = sl.HermitianMatrix(n, nb, grid).generate("rand_dominant", 42)
                Bd = sl.Matrix(n, 10, nb, grid).generate("rand", 43)
                assert sl.potrf(Hd) == 0
                sl.potrs(Hd, Bd)
                xs = gather(Bd, n, 10)
                Hm = sl.HermitianMatrix(n, nb, grid).generate("rand_dominant", 42)
                Bm = sl.Matrix(n, 10, nb, grid).generate("rand", 43)
                Xm = sl.Matrix(n, 10, nb, grid)
                minfo, mit, _ = sl.posv_mixed(Hm, Bm, Xm)
                xm = gather(Xm, n, 10)
                Gm = sl.Matrix(n, n, nb, grid).generate("rand", 42)
                Bg = sl.Matrix(n, 10, nb, grid).generate("rand", 43)
                Xg = sl.Matrix(n, 10, nb, grid)
                ginfo, git, _gp, _ = sl.gesv_mixed(Gm, Bg, Xg)
                xg = gather(Xg, n, 10)
                if rank == 0:
                    ag = o.generate("rand", n, n, 42)
                    rg = o.solve_residual(ag, xg, o.generate("rand", n, 10, 43))
                    goodg = ginfo == 0 and 0 <= git <= 30 and rg <= 25 * EPS
                    print(f"grid {p}x{q} n={n} nb={nb}: gesv_mixed iter={git} resid={rg:.2e} {'ok' if goodg else 'FAILED'}", flush=True)
                    ok &= goodg
                if rank == 0:
                    G = o.generate("rand_dominant", n, n, 42); b = o.generate("rand", n, 10, 43)
                    Lo, _ = o.potrf(o.he_full(G), nb)
                    xo = o.potrs(Lo, b, nb)
                    e1 = np.abs(xs - xo).max() / np.abs(xo).max()
                    xmo, ito, _ = o.solve_mixed(G, b, nb, hermitian=True)
                    e2 = np.abs(xm - xmo).max() / np.abs(xmo).max()
                    r2 = o.solve_residual(o.he_full(G), xm, b)
                    good = e1 <= 200 * EPS and minfo == 0 and abs(mit - ito) <= 1 and e2 <= 1e-12 and r2 <= 25 * EPS
                    print(f"grid {p}x{q} n={n} nb={nb}: potrs err={e1:.2e}; posv_mixed iter={mit} (oracle {ito}) err={e2:.2e} "
                          f"resid={r2:.2e} {'ok' if good else 'FAILED'}", flush=True)
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
