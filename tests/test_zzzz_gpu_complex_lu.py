"""GPU parity tests of the complex LU (sb200_getrf_{z,c}: cabs1 pivot rule, complex reciprocal scaling; getrf_cplx.cu)
and of the complex solve from its factors (sb200_getrs_{z,c}) through the host mirror of the reference API, against
the golden vectors of the UNMODIFIED reference (tests/golden/getrf_z.npz, getrf_c.npz, gesv_z.npz), the numpy oracle
on the same seeded inputs, and the reference tester's residual check (test/test_gesv.cc:371-377)."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o
from tests.gpu_util import GETRF_TOL

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps
CPLX_TOL = 2 * GETRF_TOL          # complex products round four times per multiply-add


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _backward_error(A0, LU, piv, nb):
    n = A0.shape[0]
    perm = o.pivots_to_perm(piv, n, nb)
    assert sorted(perm.tolist()) == list(range(n))
    L = np.tril(LU, -1) + np.eye(n)
    U = np.triu(LU)
    return np.abs(A0[perm].astype(np.complex128) - L.astype(np.complex128) @ U.astype(np.complex128)).max() / np.abs(A0).max()


def test_zgetrf_matches_reference_golden_with_identical_pivots(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "getrf_z.npz"))
    n, nb = 192, 64
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == int(g["info"]) == 0
    flat = np.array([x for c in piv for x in c], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference's"
    assert np.abs(A.to_host() - g["out"]).max() <= CPLX_TOL * np.abs(g["out"]).max()


def test_cgetrf_against_reference_golden(sl, golden_dir):
    """complex<float>: a near-tie of two candidates may resolve differently in 24-bit arithmetic, so the factor is
    compared only when the pivots agree; P A = L U has to hold either way."""
    g = np.load(os.path.join(golden_dir, "getrf_c.npz"))
    n, nb = 200, 64
    A = sl.Matrix(n, n, nb, dtype=np.complex64).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    eps32 = np.finfo(np.float32).eps
    A0 = o.generate("rand", n, n, 42, dtype=np.complex64)
    assert _backward_error(A0, LU, piv, nb) <= 64 * eps32 * n
    flat = np.array([x for c in piv for x in c], dtype=np.int64)
    if np.array_equal(flat, g["piv"]):
        assert np.abs(LU - g["out"]).max() <= 4096 * eps32 * np.abs(g["out"]).max()


@pytest.mark.parametrize("n,nb", [(700, 128), (1024, 256), (512, 512)])
def test_zgetrf_vs_oracle_and_tester_residual(sl, n, nb):
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", n, n, 42, dtype=np.complex128)
    LUo, pivo, _ = o.getrf(A0, nb, 32)
    assert piv == pivo
    assert np.abs(LU - LUo).max() <= CPLX_TOL * np.abs(LUo).max()
    assert _backward_error(A0, LU, piv, nb) <= 64 * EPS * n
    for nrhs in (10, 70):                                    # small-nrhs solve kernel / block substitution
        B = sl.Matrix(n, nrhs, nb, dtype=np.complex128).generate("rand", 43)
        sl.getrs(A, piv, B)
        B0 = o.generate("rand", n, nrhs, 43, dtype=np.complex128)
        assert o.solve_residual(A0, B.to_host(), B0) <= 50 * EPS / 2


def test_zgesv_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "gesv_z.npz"))
    n, nb, nrhs = 200, 64, 70
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=np.complex128).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == int(g["info"]) == 0
    sl.getrs(A, piv, B)
    assert np.abs(B.to_host() - g["out"]).max() <= 1e-10 * np.abs(g["out"]).max()


def test_zgetrf_rectangular_and_zero_pivot_info(sl):
    for (m, n, nb) in [(448, 256, 64), (300, 500, 128)]:
        A = sl.Matrix(m, n, nb, dtype=np.complex128).generate("rand", 7)
        piv, info = sl.getrf(A)
        A0 = o.generate("rand", m, n, 7, dtype=np.complex128)
        LUo, pivo, info_o = o.getrf(A0, nb, 32)
        assert info == info_o == 0 and piv == pivo
        assert np.abs(A.to_host() - LUo).max() <= CPLX_TOL * np.abs(LUo).max()
    n, nb = 256, 64
    A0 = o.generate("rand", n, n, 3, dtype=np.complex128); A0[:, 100] = 0.0
    A = sl.Matrix(n, n, nb, dtype=np.complex128); A.from_host(np.asfortranarray(A0))
    _, info = sl.getrf(A)
    assert info == o.getrf(A0, nb, 32)[2] == 101


@pytest.mark.parametrize("routine,kind,herm", [("gesv_mixed", "rand", False), ("posv_mixed", "rand_dominant", True)])
def test_complex_mixed_solvers_match_reference_golden(sl, golden_dir, routine, kind, herm):
    """<complex<double>, complex<float>> (src/gesv_mixed.cc:303-316): complex<float> factor, complex<double> refinement;
    the reference's iteration count (one more allowed: the FP32 roundings differ) and its solution."""
    g = np.load(os.path.join(golden_dir, routine + "_z.npz"))
    n, nb = 256, 64
    A = (sl.HermitianMatrix(n, nb, dtype=np.complex128) if herm else sl.Matrix(n, n, nb, dtype=np.complex128)).generate(kind, 42)
    B = sl.Matrix(n, 10, nb, dtype=np.complex128).generate("rand", 43)
    X = sl.Matrix(n, 10, nb, dtype=np.complex128)
    res = sl.posv_mixed(A, B, X) if herm else sl.gesv_mixed(A, B, X)
    info, it = res[0], res[1]
    assert info == int(g["info"]) == 0
    assert 0 <= it <= int(g["iters"]) + 1
    x = X.to_host()
    assert np.abs(x - g["out"]).max() <= 1e-11 * np.abs(g["out"]).max()
    a = o.generate(kind, n, n, 42, dtype=np.complex128)
    af = o.he_full(np.tril(a)) if herm else a
    assert o.solve_residual(af, x, o.generate("rand", n, 10, 43, dtype=np.complex128)) <= 25 * EPS


@pytest.mark.parametrize("routine,n,nb", [("posv", 1000, 128), ("gesv", 700, 128), ("gesv", 2048, 512)])
def test_complex_mixed_solvers_vs_oracle_and_tester_residual(sl, routine, n, nb):
    herm = routine == "posv"
    kind = "rand_dominant" if herm else "rand"
    A = (sl.HermitianMatrix(n, nb, dtype=np.complex128) if herm else sl.Matrix(n, n, nb, dtype=np.complex128)).generate(kind, 42)
    B = sl.Matrix(n, 10, nb, dtype=np.complex128).generate("rand", 43)
    X = sl.Matrix(n, 10, nb, dtype=np.complex128)
    res = sl.posv_mixed(A, B, X) if herm else sl.gesv_mixed(A, B, X)
    info, it = res[0], res[1]
    assert info == 0 and 0 <= it <= 30
    a = o.generate(kind, n, n, 42, dtype=np.complex128); b = o.generate("rand", n, 10, 43, dtype=np.complex128)
    af = o.he_full(np.tril(a)) if herm else a
    x = X.to_host()
    assert o.solve_residual(af, x, b) <= 25 * EPS
    if n <= 1000:
        xo, ito, _ = o.solve_mixed(np.tril(a) if herm else a, b, nb, hermitian=herm)
        assert abs(it - ito) <= 1
        assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()


def test_complex_mixed_not_converged_takes_the_fallback(sl):
    """itermax = 0 with a tolerance no complex<float> factor meets: iter = -(itermax + 1), complex<double> fallback solves."""
    n, nb = 300, 64
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    B = sl.Matrix(n, 4, nb, dtype=np.complex128).generate("rand", 43)
    X = sl.Matrix(n, 4, nb, dtype=np.complex128)
    info, it, piv, _ = sl.gesv_mixed(A, B, X, {"max_iterations": 0, "tolerance": 1e-30})
    assert info == 0 and it == -1
    a = o.generate("rand", n, n, 42, dtype=np.complex128); b = o.generate("rand", n, 4, 43, dtype=np.complex128)
    assert o.solve_residual(a, X.to_host(), b) <= 25 * EPS
    _, pivo, _ = o.getrf(a, nb, 32)
    assert piv == pivo


def test_complex_tntpiv_matches_reference_golden_and_tournament_oracle(sl, golden_dir, monkeypatch):
    g = np.load(os.path.join(golden_dir, "getrf_tntpiv_z.npz"))
    n, nb = 192, 64
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    piv, info = sl.lu_factor(A, {"method_lu": "CALU"})
    assert info == int(g["info"]) == 0
    assert np.array_equal(np.array([x for c in piv for x in c], dtype=np.int64), g["piv"])
    assert np.abs(A.to_host() - g["out"]).max() <= CPLX_TOL * np.abs(g["out"]).max()
    # several participants per panel (the process rows of a grid), against the restatement of internal_getrf_tntpiv.cc
    monkeypatch.setenv("SB200_TNT_RANKS", "3")
    for (m, n, nb) in [(300, 300, 64), (448, 256, 64)]:
        A = sl.Matrix(m, n, nb, dtype=np.complex128).generate("rand", 42)
        piv, info = sl.getrf_tntpiv(A)
        A0 = o.generate("rand", m, n, 42, dtype=np.complex128)
        LUo, pivo, _ = o.getrf_tntpiv(A0, nb, 32, ranks=3)
        assert info == 0 and piv == pivo
        assert np.abs(A.to_host() - LUo).max() <= CPLX_TOL * np.abs(LUo).max()


def test_complex_nopiv_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "getrf_nopiv_z.npz"))
    n, nb = 200, 64
    A = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand_dominant", 42)
    assert sl.getrf_nopiv(A) == int(g["info"]) == 0
    assert np.abs(A.to_host() - g["out"]).max() <= CPLX_TOL * np.abs(g["out"]).max()
    # complex<float>: A = L U to working precision; the pivot search is back on the next call
    C = sl.Matrix(n, n, nb, dtype=np.complex64).generate("rand_dominant", 42)
    assert sl.getrf_nopiv(C) == 0
    LU = C.to_host().astype(np.complex128)
    A0 = o.generate("rand_dominant", n, n, 42, dtype=np.complex64).astype(np.complex128)
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    assert np.abs(L @ U - A0).max() <= 64 * np.finfo(np.float32).eps * np.abs(A0).max()
    B = sl.Matrix(n, n, nb, dtype=np.complex128).generate("rand", 42)
    piv, _ = sl.getrf(B)
    assert piv == o.getrf(o.generate("rand", n, n, 42, dtype=np.complex128), nb, 32)[1]
