"""GPU tests of the p x q solve path (slate_b200/csrc/solve_dist.cu: replicated right-hand sides) through the single-rank
test hook SB200_DIST_SOLVE=2 (same code path as on a grid, minus the NCCL calls).

Validated on a B200 in round 2 (gpurun call r2b: all green; profiles/r02b_pytest_gpu_tail.txt).  The multi-rank check
itself lives in scratch/mgpu_check.py (potrs / posv_mixed on 1x2 and 2x1 grids)."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o

pytestmark = [pytest.mark.gpu]
EPS = float(np.finfo(np.float64).eps)


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


@pytest.mark.parametrize("n,nb,nrhs", [(1024, 256, 10), (1000, 128, 5), (300, 512, 3)])
def test_potrs_replicated_rhs_path(sl, monkeypatch, n, nb, nrhs):
    monkeypatch.setenv("SB200_DIST_SOLVE", "2")
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb).generate("rand", 43)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    sl.potrs(A, B)
    Xo = o.potrs(L, o.generate("rand", n, nrhs, 43), nb)
    assert np.abs(B.to_host() - Xo).max() <= 200 * EPS * np.abs(Xo).max()


@pytest.mark.parametrize("n,nb,nrhs", [(1024, 256, 10), (700, 128, 7)])
def test_getrs_replicated_rhs_path(sl, monkeypatch, n, nb, nrhs):
    monkeypatch.setenv("SB200_DIST_SOLVE", "2")
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    sl.getrs(A, piv, B)
    X = B.to_host()
    assert o.solve_residual(o.generate("rand", n, n, 42), X, o.generate("rand", n, nrhs, 43)) <= 25 * EPS


@pytest.mark.parametrize("routine,n,nb", [("posv", 1024, 256), ("gesv", 1024, 256), ("posv", 1000, 128), ("gesv", 700, 128)])
def test_mixed_solvers_replicated_rhs_path(sl, monkeypatch, routine, n, nb):
    monkeypatch.setenv("SB200_DIST_SOLVE", "2")
    monkeypatch.setenv("SB200_GETRF_DIST", "1")
    herm = routine == "posv"
    kind = "rand_dominant" if herm else "rand"
    A = (sl.HermitianMatrix(n, nb) if herm else sl.Matrix(n, n, nb)).generate(kind, 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    res = sl.posv_mixed(A, B, X) if herm else sl.gesv_mixed(A, B, X)
    a = o.generate(kind, n, n, 42); b = o.generate("rand", n, 10, 43)
    xo, ito, _ = o.solve_mixed(a, b, nb, hermitian=herm)
    assert res[0] == 0 and abs(res[1] - ito) <= 1
    x = X.to_host()
    assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
    assert o.solve_residual(o.he_full(a) if herm else a, x, b) <= 25 * EPS


def test_potrf_streams_finished_columns_to_host(sl):
    """potrf(out_local=...) (D2H of every block column as soon as it is final, overlapped with the trailing updates)
    must deliver exactly what to_host_local delivers afterwards."""
    import torch
    n, nb = 2048, 256
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    out = torch.empty(A.local_tiles * nb * nb, dtype=torch.float64).pin_memory()
    ref = torch.empty_like(out)
    assert sl.potrf(A, out_local=out) == 0
    A.to_host_local(ref)
    assert torch.equal(out, ref)
