"""GPU parity tests of getrf_tntpiv (LU with tournament pivoting, CALU) through the host
mirror of the reference API, against
  (1) the golden vectors the UNMODIFIED reference wrote with Option::MethodLU = CALU (tests/golden/getrf_tntpiv_d*.npz;
      one MPI rank: the only case the reference can run where there is no MPI),
  (2) the numpy restatement of src/internal/internal_getrf_tntpiv.cc on the same seeded inputs, for one and for several
      participants per panel (SB200_TNT_RANKS = the process rows of a grid: the tournament runs on the GPU that holds the
      gathered panel, so a 1 x 1 grid executes exactly the code a p x q grid executes after its gather),
  (3) P A = L U and the reference tester's solve residual."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o
from tests.gpu_util import GETRF_TOL

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _check_lu(A0, LU, piv, nb):
    m, n = A0.shape
    mn = min(m, n)
    perm = o.pivots_to_perm(piv, m, nb)
    assert sorted(perm.tolist()) == list(range(m))
    L = np.tril(LU, -1)[:, :mn] + np.eye(m, mn)
    U = np.triu(LU)[:mn]
    assert np.abs(A0[perm] - L @ U).max() <= 64 * EPS * max(m, n) * np.abs(A0).max()
    return perm, L, U


@pytest.mark.parametrize("name,m,n", [("getrf_tntpiv_d", 384, 384), ("getrf_tntpiv_d_ragged", 300, 300),
                                      ("getrf_tntpiv_d_tall", 512, 256)])
def test_tntpiv_matches_reference_golden_with_identical_pivots(sl, golden_dir, name, m, n):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    A = sl.Matrix(m, n, 128).generate("rand", 42)
    piv, info = sl.getrf_tntpiv(A)
    assert info == int(g["info"]) == 0
    flat = np.array([x for c in piv for x in c], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference's"
    LU = A.to_host()
    assert np.abs(LU - g["out"]).max() <= GETRF_TOL * np.abs(g["out"]).max()


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("n,nb", [(1024, 256), (700, 128)])
def test_tntpiv_one_participant_vs_oracle_and_tester_residual(sl, t, n, nb):
    dt = np.float64 if t == "d" else np.float32
    A = sl.Matrix(n, n, nb, dtype=dt).generate("rand", 42)
    piv, info = sl.lu_factor(A, {"method_lu": "CALU"})
    assert info == 0
    LU = A.to_host().astype(np.float64)
    A0 = o.generate("rand", n, n, 42, dtype=dt)
    LUo, pivo, _ = o.getrf_tntpiv(A0, nb, 32)
    if t == "d":
        assert piv == pivo
        assert np.abs(LU - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
        perm, L, U = _check_lu(A0, LU, piv, nb)
        B = o.generate("rand", n, 10, 43)
        X = np.linalg.solve(U, np.linalg.solve(L, B[perm]))
        assert o.solve_residual(A0, X, B) <= 50 * EPS / 2          # test/test_gesv.cc:371-377
    else:
        eps32 = np.finfo(np.float32).eps
        perm = o.pivots_to_perm(piv, n, nb)
        L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
        assert np.abs(A0.astype(np.float64)[perm] - L @ U).max() <= 64 * eps32 * n


@pytest.mark.parametrize("ranks", [2, 3, 4])
@pytest.mark.parametrize("m,n,nb", [(384, 384, 64), (300, 300, 64), (448, 256, 64), (1024, 1024, 128)])
def test_tntpiv_tournament_matches_oracle(sl, monkeypatch, ranks, m, n, nb):
    monkeypatch.setenv("SB200_TNT_RANKS", str(ranks))
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf_tntpiv(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, _ = o.getrf_tntpiv(A0, nb, 32, ranks=ranks)
    assert piv == pivo, "the tournament picked other rows than the restatement of internal_getrf_tntpiv.cc"
    assert np.abs(LU - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
    _check_lu(A0, LU, piv, nb)


@pytest.mark.parametrize("name,ranks,m,n", [("grid_getrf_tntpiv_d_2x1", 2, 384, 384), ("grid_getrf_tntpiv_d_3x1_ragged", 3, 300, 300),
                                            ("grid_getrf_tntpiv_d_4x1_tall", 4, 448, 256)])
def test_tntpiv_tournament_matches_multirank_reference_golden(sl, golden_dir, monkeypatch, name, ranks, m, n):
    """Against what the UNMODIFIED reference computed on `ranks` x 1 process grids (tests/golden/grid_*.npz, written through
    oracle/mprun.py): identical pivots, factor within the LU bound.  (Added after the round's GPU budget ended: the same
    cases are GPU-validated against the oracle above, and the oracle is pinned to these files in tests/test_oracle.py.)"""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    monkeypatch.setenv("SB200_TNT_RANKS", str(ranks))
    A = sl.Matrix(m, n, 64).generate("rand", 42)
    piv, info = sl.getrf_tntpiv(A)
    assert info == int(g["info"]) == 0
    assert np.array_equal(np.array([x for c in piv for x in c], dtype=np.int64), g["piv"])
    assert np.abs(A.to_host() - g["out"]).max() <= GETRF_TOL * np.abs(g["out"]).max()


@pytest.mark.parametrize("ranks", [1, 2])
@pytest.mark.parametrize("m,n,nb", [(384, 384, 128), (300, 300, 64), (448, 256, 64)])
def test_tntpiv_grid_algorithm_on_one_rank(sl, monkeypatch, ranks, m, n, nb):
    """The p x q driver (getrf_dist.cu) with the tournament panel: row map, permutation by gather / scatter, U workspace."""
    monkeypatch.setenv("SB200_GETRF_DIST", "1")
    monkeypatch.setenv("SB200_TNT_RANKS", str(ranks))
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf_tntpiv(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, _ = o.getrf_tntpiv(A0, nb, 32, ranks=ranks)
    assert piv == pivo
    assert np.abs(LU - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
    _check_lu(A0, LU, piv, nb)


def test_tntpiv_rejects_the_shapes_the_reference_rejects(sl):
    # a diagonal tile that is not square: the reference throws (include/slate/TriangularMatrix.hh:459)
    for (m, n, nb) in [(500, 300, 128), (300, 500, 128)]:
        A = sl.Matrix(m, n, nb).generate("rand", 42)
        with pytest.raises(sl.SB200Error):
            sl.getrf_tntpiv(A)
    with pytest.raises(sl.SB200Error):
        sl.lu_factor(sl.Matrix(128, 128, 64).generate("rand", 1), {"method_lu": "no-such-method"})


def test_lu_factor_method_dispatch(sl):
    n, nb = 256, 64
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, n, nb).generate("rand", 42)
    p1, _ = sl.lu_factor(A)
    p2, _ = sl.getrf(B)
    assert p1 == p2 and np.array_equal(A.to_host(), B.to_host())
    D = sl.Matrix(n, n, nb).generate("rand_dominant", 42)
    piv, info = sl.lu_factor(D, {"method_lu": "NoPiv"})
    assert piv == [] and info == 0
    LUo, _ = o.getrf_nopiv(o.generate("rand_dominant", n, n, 42), nb)
    assert np.abs(D.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
