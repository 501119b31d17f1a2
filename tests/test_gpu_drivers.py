"""GPU parity tests of the drivers (pytest -m gpu): potrf / getrf / gemm through the host mirror
of the reference API (slate_b200.host) against (1) the golden vectors written by the UNMODIFIED
reference (tests/golden), (2) the numpy oracle on the same seeded inputs, (3) the reference
tester's residual checks, and (4) size-independent properties at BASELINE.json's full size."""
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


def test_device_generator_is_bit_exact(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "gen_d.npz"))
    A = sl.Matrix(80, 96, 32).generate("rand", 5)
    assert np.array_equal(A.to_host(), g["rand"])
    H = sl.HermitianMatrix(96, 32).generate("rand_dominant", 7)
    assert np.array_equal(np.tril(H.to_host()), np.tril(g["rand_dominant"]))


def test_potrf_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "potrf_d.npz"))
    A = sl.HermitianMatrix(384, 128).generate("rand_dominant", 42)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    ref = np.tril(g["out"])
    assert np.abs(L - ref).max() <= 32 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("n,nb", [(512, 128), (1000, 128), (1536, 512), (2048, 256), (96, 128)])
def test_potrf_vs_oracle_and_tester_residual(sl, n, nb):
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    G = o.generate("rand_dominant", n, n, 42)
    Af = np.tril(G) + np.tril(G, -1).T
    Lo, info = o.potrf(Af, nb)
    assert info == 0
    assert np.abs(L - Lo).max() <= 64 * EPS * np.abs(Lo).max()
    B = o.generate("rand", n, 10, 43)
    X = np.linalg.solve(L.T, np.linalg.solve(L, B))
    assert o.solve_residual(Af, X, B) <= 50 * EPS / 2          # test/test_posv.cc:336-342


def test_potrf_info_matches_oracle(sl):
    n, nb = 512, 128
    H = o.generate("rand", n, n, 1); H = H + H.T + n * np.eye(n); H[300, 300] = -5.0
    A = sl.HermitianMatrix(n, nb); A.from_host(np.asfortranarray(H))
    _, info_ref = o.potrf(H, nb)
    assert sl.potrf(A) == info_ref == 301


@pytest.mark.parametrize("name,n", [("getrf_d", 384), ("getrf_d_ragged", 300)])
def test_getrf_matches_reference_golden_with_identical_pivots(sl, golden_dir, name, n):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    A = sl.Matrix(n, n, 128).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    flat = np.array([x for c in piv for x in c], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference's"
    LU = A.to_host()
    assert np.abs(LU - g["out"]).max() <= GETRF_TOL * np.abs(g["out"]).max()


@pytest.mark.parametrize("n,nb", [(512, 512), (1024, 256), (2048, 512), (700, 128)])
def test_getrf_vs_oracle_and_tester_residual(sl, n, nb):
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", n, n, 42)
    LUo, pivo, _ = o.getrf(A0, nb, 32)
    assert piv == pivo
    assert np.abs(LU - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
    perm = o.pivots_to_perm(piv, n, nb)
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    B = o.generate("rand", n, 10, 43)
    X = np.linalg.solve(U, np.linalg.solve(L, B[perm]))
    assert o.solve_residual(A0, X, B) <= 50 * EPS / 2          # test/test_gesv.cc:371-377


def test_getrf_zero_pivot_info(sl):
    n, nb = 256, 64
    A0 = o.generate("rand", n, n, 3); A0[:, 100] = 0.0
    A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
    _, info = sl.getrf(A)
    _, _, info_ref = o.getrf(A0, nb, 32)
    assert info == info_ref == 101


def test_gemm_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "gemm_d.npz"))
    n, nb = 256, 64
    A = sl.Matrix(n, n, nb).generate("rand", 42); B = sl.Matrix(n, n, nb).generate("rand", 43)
    C = sl.Matrix(n, n, nb).generate("rand", 44)
    al, be = 3.141592653589793, 2.718281828459045
    sl.gemm(al, A, B, be, C)
    out = C.to_host()
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    a, b, c0 = (o.generate("rand", n, n, s) for s in (42, 43, 44))
    assert o.gemm_check(al, a, b, be, c0, out) <= 3 * EPS       # test/test_gemm.cc:205-207


@pytest.mark.parametrize("m,n,k,nb", [(300, 200, 150, 64), (1024, 1024, 1024, 256)])
def test_gemm_rectangular_ragged(sl, m, n, k, nb):
    A = sl.Matrix(m, k, nb).generate("rand", 1); B = sl.Matrix(k, n, nb).generate("rand", 2)
    C = sl.Matrix(m, n, nb).generate("rand", 3)
    sl.gemm(-1.0, A, B, 0.5, C)
    a = o.generate("rand", m, k, 1); b = o.generate("rand", k, n, 2); c = o.generate("rand", m, n, 3)
    ref = o.gemm(-1.0, a, b, 0.5, c, nb)
    assert np.abs(C.to_host() - ref).max() <= 64 * EPS * np.abs(ref).max()


def test_host_round_trip_and_errors(sl):
    n, nb = 200, 64
    H = np.asfortranarray(o.generate("rand", n, n, 9))
    A = sl.Matrix(n, n, nb); A.from_host(H)
    assert np.array_equal(A.to_host(), H)
    with pytest.raises(sl.Exception_):
        A.from_host(np.zeros((n, n)))                          # C-ordered host array is rejected
    with pytest.raises(Exception):
        sl.potrf(A)                                            # potrf on a general matrix: invalid argument


def test_full_size_potrf_properties(sl):
    """BASELINE configs[1]: dpotrf n=32768 nb=512 -- checked through size-independent properties
    (the oracle would need minutes): info == 0, and the tester residual evaluated on the device-
    generated data with a probe-vector identity  ||A x - L (L^T x)|| / (n ||A|| ||x||)."""
    import torch
    n, nb = 32768, 512
    free, _ = torch.cuda.mem_get_info()
    if free < 24 * 2 ** 30:
        pytest.skip("not enough free HBM for the full-size case")
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    host = torch.empty((n, n), dtype=torch.float64)
    A.to_host(host)
    G = host.numpy().T                                         # column-major view of the lower tiles
    x = o.generate("rand", n, 1, 5)[:, 0]
    Gl = np.tril(G)
    ax = Gl @ x + np.tril(G, -1).T @ x
    a_norm1 = float(np.abs(Gl).sum(axis=0).max()) * 2          # upper bound of ||A||_1
    del Gl
    assert sl.potrf(A) == 0
    A.to_host(host)
    L = np.tril(host.numpy().T)
    llx = L @ (L.T @ x)
    res = np.abs(ax - llx).sum() / (n * a_norm1 * np.abs(x).sum())
    assert res <= 50 * EPS / 2
    assert A.last_driver_ms > 0


@pytest.mark.parametrize("n,nb", [(384, 128), (300, 128), (1024, 256), (2048, 512)])
def test_getrf_grid_algorithm_on_one_rank(sl, n, nb, monkeypatch):
    """The p x q LU driver (getrf_dist.cu: panel workspace, row-map permutation instead of sequential
    swaps, U workspace) forced onto a 1 x 1 grid: must reproduce the oracle's pivots and factors."""
    monkeypatch.setenv("SB200_GETRF_DIST", "1")
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", n, n, 42)
    LUo, pivo, _ = o.getrf(A0, nb, 32)
    assert piv == pivo
    assert np.abs(LU - LUo).max() <= GETRF_TOL * np.abs(LUo).max()


def test_getrf_grid_algorithm_rectangular_and_singular(sl, monkeypatch):
    monkeypatch.setenv("SB200_GETRF_DIST", "1")
    for (m, n, nb) in [(700, 300, 128), (300, 700, 128)]:
        A = sl.Matrix(m, n, nb).generate("rand", 7)
        piv, info = sl.getrf(A)
        A0 = o.generate("rand", m, n, 7)
        LUo, pivo, info_o = o.getrf(A0, nb, 32)
        assert info == info_o == 0 and piv == pivo
        assert np.abs(A.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()
    n, nb = 256, 64
    A0 = o.generate("rand", n, n, 3); A0[:, 100] = 0.0
    A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
    _, info = sl.getrf(A)
    assert info == o.getrf(A0, nb, 32)[2] == 101


def test_bcast_tiles_hook_on_one_rank_is_the_single_rank_early_exit(sl):
    """sb200_bcast_tiles on a 1 x 1 grid returns at once and touches nothing (BaseMatrix.hh:2006: one rank, no
    broadcast); the multi-rank transfer itself is checked by scratch/mgpu_check.py on 2 and 8 GPUs."""
    import torch
    g = sl.Grid(1, 1, 0)
    a = torch.arange(1024, dtype=torch.float64, device="cuda")
    b = torch.zeros(1024, dtype=torch.float64, device="cuda")
    g.bcast_tiles([(a.data_ptr(), b.data_ptr(), 8192, 0)])
    torch.cuda.synchronize()
    assert float(b.abs().sum()) == 0.0
    g.close()


def test_gemm_baseline_config0_full_size(sl):
    """BASELINE.json configs[0] at its full size: dgemm n = 4096, nb = 256, tester alpha / beta and seeds
    (test/test.cc:447-448), against the FP64 product of the same inputs and the tester's check (test/test_gemm.cc:205-207)."""
    n, nb = 4096, 256
    A = sl.Matrix(n, n, nb).generate("rand", 42); B = sl.Matrix(n, n, nb).generate("rand", 43)
    C = sl.Matrix(n, n, nb).generate("rand", 44)
    al, be = 3.141592653589793, 2.718281828459045
    # inputs read back from the device: its Philox generator is pinned bit-exact to the reference's matgen by
    # test_device_generator_is_bit_exact (the numpy generator would take half a minute at this size)
    a, b, c0 = A.to_host(), B.to_host(), C.to_host()
    sl.gemm(al, A, B, be, C)
    out = C.to_host()
    ref = al * (a @ b) + be * c0
    assert np.abs(out - ref).max() <= 3 * np.sqrt(n) * EPS * np.abs(ref).max()       # the tester's 3 sqrt(k) eps scale
    assert o.gemm_check(al, a, b, be, c0, out) <= 3 * EPS
