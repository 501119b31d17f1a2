"""GPU tests (pytest -m gpu) of the probe-vector products and the bench-size residual checks built on them
(slate_b200/csrc/probe.cu, host.potrf_residual / getrf_residual / gemm_residual): each part / op against numpy on the
same seeded matrix, then the residuals of correct and of deliberately damaged factors."""
import numpy as np
import pytest

from oracle import slate_oracle as o

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _vec(n, dtype, seed):
    import torch
    rng = np.random.default_rng(seed)
    v = rng.random(n) - 0.5
    if np.dtype(dtype).kind == "c":
        v = v + 1j * (rng.random(n) - 0.5)
    return torch.from_numpy(v.astype(dtype)).cuda(), v.astype(dtype)


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
@pytest.mark.parametrize("m,n,nb", [(300, 200, 64), (512, 512, 128), (100, 260, 128)])
def test_probe_mv_general_parts(sl, t, m, n, nb):
    A = sl.Matrix(m, n, nb, dtype=t).generate("rand", 11)
    Ah = A.to_host()
    tol = 50 * np.finfo(Ah.real.dtype).eps * max(m, n)
    x, xh = _vec(n, Ah.dtype, 1)
    xm, xmh = _vec(m, Ah.dtype, 2)
    for part, M in (("G", Ah), ("L", np.tril(Ah)), ("U", np.triu(Ah))):
        y = sl.probe_mv(A, x, part, "N").cpu().numpy()
        assert np.abs(y - M @ xh).max() <= tol, (part, "N")
        y = sl.probe_mv(A, xm, part, "C").cpu().numpy()
        assert np.abs(y - M.conj().T @ xmh).max() <= tol, (part, "C")
    Lu = np.tril(Ah, -1) + np.eye(m, n, dtype=Ah.dtype)
    y = sl.probe_mv(A, x, "L", "N", "U").cpu().numpy()
    assert np.abs(y - Lu @ xh).max() <= tol
    ones = x.new_ones(m)
    cs = sl.probe_mv(A, ones, "G", "C", use_abs=True).cpu().numpy()
    assert np.abs(cs.real - np.abs(Ah).sum(axis=0)).max() <= tol


@pytest.mark.parametrize("t", ["d", "z"])
@pytest.mark.parametrize("n,nb", [(384, 128), (300, 128)])
def test_probe_mv_hermitian(sl, t, n, nb):
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand_dominant", 42)
    Lh = np.tril(A.to_host())
    H = Lh + np.tril(Lh, -1).conj().T
    H[np.diag_indices(n)] = H[np.diag_indices(n)].real
    x, xh = _vec(n, Lh.dtype, 3)
    y = sl.probe_mv(A, x, "H").cpu().numpy()
    assert np.abs(y - H @ xh).max() <= 50 * EPS * n * np.abs(H).max()
    rs = sl.probe_mv(A, x.new_ones(n), "H", use_abs=True).cpu().numpy()
    assert np.abs(rs.real - np.abs(H).sum(axis=1)).max() <= 50 * EPS * n * np.abs(H).max()


@pytest.mark.parametrize("n,nb", [(1024, 256), (1000, 128)])
def test_potrf_residual_passes_and_detects_damage(sl, n, nb):
    A0 = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    A = sl.HermitianMatrix(n, nb); A.copy_from(A0)
    assert sl.potrf(A) == 0
    r = sl.potrf_residual(A0, A)
    assert r["pass"] and r["residual"] <= 25 * EPS, r
    L = A.to_host(); L[n // 2, n // 3] += 1e-6 * abs(L[n // 2, n // 3]) + 1e-6
    A.from_host(np.asfortranarray(L))
    assert not sl.potrf_residual(A0, A)["pass"]


@pytest.mark.parametrize("m,n,nb", [(1024, 1024, 256), (700, 700, 128), (900, 500, 128)])
def test_getrf_residual_passes_and_detects_damage(sl, m, n, nb):
    A0 = sl.Matrix(m, n, nb).generate("rand", 42)
    A = sl.Matrix(m, n, nb); A.copy_from(A0)
    piv, info = sl.getrf(A)
    assert info == 0
    r = sl.getrf_residual(A0, A, piv)
    assert r["pass"], r
    bad = [list(b) for b in piv]
    j = next(i for i, (t, off) in enumerate(bad[0]) if (t, off) != (0, i))      # undo one real interchange
    bad[0][j] = (0, j)
    assert not sl.getrf_residual(A0, A, bad)["pass"]


def test_gemm_residual(sl):
    m, n, k, nb = 700, 500, 900, 128
    A = sl.Matrix(m, k, nb).generate("rand", 1)
    B = sl.Matrix(k, n, nb).generate("rand", 2)
    C = sl.Matrix(m, n, nb).generate("rand", 3)
    x = sl._probe_vector(n, C.dtype, 5)
    c0x = sl.probe_mv(C, x, "G")
    sl.gemm(3.1, A, B, 2.7, C)
    assert sl.gemm_residual(3.1, A, B, 2.7, c0x, C, x)["pass"]
    assert not sl.gemm_residual(3.1, A, B, 2.5, c0x, C, x)["pass"]
