"""GPU parity tests (pytest -m gpu) of the two LU panel algorithms and of the fast diagonal-block
kernels: both SB200_PANEL=1 (right-looking 32-column blocks + laswp launches) and SB200_PANEL=2
(recursive panel, panel-wide interchanges inside the cooperative kernel) must give the pivot
vectors of the reference (oracle restatement of src/internal/Tile_getrf.hh:196-289) and factors
within 1e-11 of it; the single-rank and the p x q drivers are both run."""
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


@pytest.mark.parametrize("dist", ["0", "1"])
@pytest.mark.parametrize("panel", ["1", "2"])
@pytest.mark.parametrize("m,n,nb", [(1024, 1024, 256), (700, 700, 128), (2048, 2048, 512),
                                    (700, 300, 128), (300, 700, 128), (1100, 1100, 512)])
def test_getrf_panel_variants_identical_pivots(sl, m, n, nb, panel, dist, monkeypatch):
    monkeypatch.setenv("SB200_PANEL", panel)
    monkeypatch.setenv("SB200_GETRF_DIST", dist)
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, info_o = o.getrf(A0, nb, 32)
    assert info == info_o == 0
    assert piv == pivo, "pivot vectors differ from the oracle's"
    assert np.abs(A.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()


@pytest.mark.parametrize("panel", ["1", "2"])
def test_getrf_panel_variants_zero_column(sl, panel, monkeypatch):
    monkeypatch.setenv("SB200_PANEL", panel)
    n, nb = 256, 64
    A0 = o.generate("rand", n, n, 3); A0[:, 100] = 0.0
    A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
    _, info = sl.getrf(A)
    assert info == o.getrf(A0, nb, 32)[2] == 101


def test_getrf_tall_panel_many_ctas(sl):
    """m_p = 8192 rows: the cooperative base kernel runs on 11 CTAs + the interchange CTA; checked
    with the reference tester's solve residual (test/test_gesv.cc:371-377) and the growth-free
    identity P A = L U on the seeded matrix."""
    n, nb = 8192, 512
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    A0 = o.generate("rand", n, n, 42)
    perm = o.pivots_to_perm(piv, n, nb)
    assert sorted(perm) == list(range(n))
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    assert np.abs(L).max() <= 1.0 + 1e-14                      # partial pivoting: |l_ij| <= 1
    x = o.generate("rand", n, 1, 5)[:, 0]
    lhs = A0[perm] @ x
    rhs = L @ (U @ x)
    assert np.abs(lhs - rhs).max() <= 1e-10 * np.abs(lhs).max()


@pytest.mark.parametrize("n,nb", [(512, 512), (448, 512), (1000, 256), (70, 128)])
def test_potrf_fast_diag_kernel_vs_oracle(sl, n, nb):
    """potrf_diag_fast_kernel / trtri_diag_fast_kernel (64-thread register kernels) through the driver:
    ragged 64-blocks (n % 64 != 0) included."""
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 11)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    G = o.generate("rand_dominant", n, n, 11)
    Af = np.tril(G) + np.tril(G, -1).T
    Lo, info = o.potrf(Af, nb)
    assert info == 0
    assert np.abs(L - Lo).max() <= 64 * EPS * np.abs(Lo).max()
    assert np.abs(L @ L.T - Af).max() <= 64 * EPS * np.abs(Af).max()
