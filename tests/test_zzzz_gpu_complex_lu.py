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
