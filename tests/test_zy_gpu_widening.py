"""GPU parity tests (pytest -m gpu) of the round-1 widening of the runtime, through slate_b200.host -> C ABI:
all four scalar types (generator, gemm, herk, potrf), the solve path (potrs, getrs, hemm, inf-norm), the FP32
factorisations on the tcgen05 FP32-emulated kernel, and posv_mixed / gesv_mixed.

Checked against (1) golden vectors written by the UNMODIFIED reference (tests/golden), (2) the numpy oracle on
the same seeded inputs, (3) the reference tester's residual checks.  Tolerances are stated per test:
FP64 paths a few hundred eps64; FP32 paths c * n * eps32 (backward-error class of an FP32 factorisation)."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float64).eps)
EPS32 = float(np.finfo(np.float32).eps)
NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
TOL = {"s": EPS32, "d": EPS, "c": EPS32, "z": EPS}
ALPHA = 3.141592653589793 + 1.414213562373095j     # tester defaults (test/test.cc:447-448)
BETA = 2.718281828459045 + 1.732050807568877j


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _sc(t, v):
    return v if t in "cz" else v.real


@pytest.mark.parametrize("t", ["s", "c", "z"])
def test_device_generator_bit_exact_all_types(sl, golden_dir, t):
    g = np.load(os.path.join(golden_dir, f"gen_{t}.npz"))
    A = sl.Matrix(80, 96, 32, dtype=t).generate("rand", 5)
    assert np.array_equal(A.to_host(), g["rand"])
    H = sl.HermitianMatrix(96, 32, dtype=t).generate("rand_dominant", 7)
    assert np.array_equal(np.tril(H.to_host()), np.tril(g["rand_dominant"]))


def test_gemm_z_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "gemm_z.npz"))["out"]
    n, nb = 256, 64
    A = sl.Matrix(n, n, nb, dtype="z").generate("rand", 42)
    B = sl.Matrix(n, n, nb, dtype="z").generate("rand", 43)
    C = sl.Matrix(n, n, nb, dtype="z").generate("rand", 44)
    sl.gemm(ALPHA, A, B, BETA, C)
    assert np.abs(C.to_host() - g).max() <= 3 * np.sqrt(n) * EPS * 4 * np.abs(g).max()


@pytest.mark.parametrize("t,m,n,k,nb", [("z", 300, 200, 260, 64), ("s", 256, 256, 256, 128), ("c", 200, 136, 72, 64),
                                        ("z", 1024, 1024, 1024, 512)])
def test_gemm_driver_all_types_vs_oracle(sl, t, m, n, k, nb):
    A = sl.Matrix(m, k, nb, dtype=t).generate("rand", 1)
    B = sl.Matrix(k, n, nb, dtype=t).generate("rand", 2)
    C = sl.Matrix(m, n, nb, dtype=t).generate("rand", 3)
    al, be = _sc(t, ALPHA), _sc(t, BETA)
    sl.gemm(al, A, B, be, C)
    a, b, c0 = (o.generate("rand", *shp, sd, NP[t]) for shp, sd in (((m, k), 1), ((k, n), 2), ((m, n), 3)))
    ref = o.gemm(al, a.astype(np.complex128 if t in "cz" else np.float64), b.astype(np.complex128 if t in "cz" else np.float64),
                 be, c0.astype(np.complex128 if t in "cz" else np.float64), nb)
    # unit_test/test_internal_blas.cc:283-452: 3 sqrt(k) eps (x4: complex products and alpha/beta scaling)
    assert np.abs(C.to_host() - ref).max() <= 3 * np.sqrt(k) * TOL[t] * 4 * np.abs(ref).max()


@pytest.mark.parametrize("t", ["d", "z"])
def test_herk_driver_matches_reference_golden(sl, golden_dir, t):
    g = np.load(os.path.join(golden_dir, f"herk_{t}.npz"))["out"]
    n, k, nb = 256, 128, 64
    A = sl.Matrix(n, k, nb, dtype=t).generate("rand", 42)
    C = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 44)
    sl.herk(ALPHA.real, A, BETA.real, C)
    out = np.tril(C.to_host())
    assert np.abs(out - np.tril(g)).max() <= 3 * np.sqrt(k) * EPS * 4 * np.abs(g).max()


@pytest.mark.parametrize("t,n,k,nb", [("z", 300, 200, 64), ("s", 256, 192, 128), ("c", 136, 72, 64), ("d", 1000, 600, 256)])
def test_herk_driver_all_types_vs_oracle(sl, t, n, k, nb):
    A = sl.Matrix(n, k, nb, dtype=t).generate("rand", 5)
    C = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 6)
    c_before_upper = np.triu(C.to_host(), 1)
    sl.herk(-1.0, A, 2.0, C)
    hi = np.complex128 if t in "cz" else np.float64
    ref = o.herk(-1.0, o.generate("rand", n, k, 5, NP[t]).astype(hi), 2.0, o.generate("rand", n, n, 6, NP[t]).astype(hi), nb)
    out = C.to_host()
    assert np.abs(np.tril(out) - np.tril(ref)).max() <= 3 * np.sqrt(k) * TOL[t] * 4 * np.abs(ref).max()
    # strictly upper tiles are not stored; the upper triangle INSIDE diagonal tiles must be untouched
    up = np.triu(out, 1)
    for s0 in range(0, n, nb):
        s1 = min(s0 + nb, n)
        assert np.array_equal(up[s0:s1, s0:s1], c_before_upper[s0:s1, s0:s1])
    if t in "cz":
        assert np.all(np.diag(out).imag == 0)            # herk leaves a real diagonal (as cublasZherk)


def test_potrf_z_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "potrf_z.npz"))
    A = sl.HermitianMatrix(192, 64, dtype="z").generate("rand_dominant", 42)
    assert sl.potrf(A) == int(g["info"]) == 0
    L = np.tril(A.to_host())
    assert np.abs(L - np.tril(g["out"])).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("t,n,nb", [("z", 520, 128), ("c", 300, 64), ("s", 1024, 256)])
def test_potrf_all_types_vs_oracle(sl, t, n, nb):
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand_dominant", 42)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    hi = np.complex128 if t in "cz" else np.float64
    G = o.generate("rand_dominant", n, n, 42, NP[t]).astype(hi)
    Lo, info = o.potrf(o.he_full(G), nb)
    assert info == 0
    assert np.abs(L - Lo).max() <= 64 * TOL[t] * np.abs(Lo).max()


@pytest.mark.parametrize("n,nb", [(1024, 256), (1536, 512), (1000, 128), (300, 512)])
def test_potrf_float_tensor_core_path(sl, n, nb):
    """FP32 Cholesky whose trailing update runs on the tcgen05 3xTF32 kernel: same accuracy class as the FP32
    SIMT path (both compared with the FP64 factor of the same float matrix), and the tester residual at FP32."""
    G = o.generate("rand_dominant", n, n, 42, np.float32).astype(np.float64)
    Lo, _ = o.potrf(o.he_full(G), nb)
    errs = {}
    for tc in (False, True):
        A = sl.HermitianMatrix(n, nb, dtype="s").generate("rand_dominant", 42)
        before_upper = np.triu(A.to_host(), 1)
        assert sl.potrf(A, {"tensor_core_fp32": tc}) == 0
        out = A.to_host()
        L = np.tril(out).astype(np.float64)
        errs[tc] = np.abs(L - Lo).max() / np.abs(Lo).max()
        assert errs[tc] <= 64 * EPS32
        for s0 in range(0, n, nb):                          # triangle mask of the diagonal tiles
            s1 = min(s0 + nb, n)
            assert np.array_equal(np.triu(out, 1)[s0:s1, s0:s1], before_upper[s0:s1, s0:s1])
        resid = np.abs(L @ L.T - o.he_full(G)).max() / (n * np.abs(G).max())
        assert resid <= 3 * EPS32
    assert errs[True] <= 8 * errs[False] + 4 * EPS32


def test_potrs_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "posv_d.npz"))
    n, nb = 300, 128
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    assert sl.potrf(A) == 0
    sl.potrs(A, B)
    X = B.to_host()
    assert np.abs(X - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    G = o.generate("rand_dominant", n, n, 42)
    assert o.solve_residual(o.he_full(G), X, o.generate("rand", n, 10, 43)) <= 25 * EPS


def test_potrs_z_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "posv_z.npz"))
    n, nb, nrhs = 192, 64, 70
    A = sl.HermitianMatrix(n, nb, dtype="z").generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb, dtype="z").generate("rand", 43)
    assert sl.potrf(A) == 0
    sl.potrs(A, B)
    assert np.abs(B.to_host() - g["out"]).max() <= 256 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("n,nb,nrhs", [(1024, 256, 10), (1000, 128, 300), (2048, 512, 1)])
def test_potrs_vs_oracle(sl, n, nb, nrhs):
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb).generate("rand", 43)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    sl.potrs(A, B)
    Xo = o.potrs(L, o.generate("rand", n, nrhs, 43), nb)
    assert np.abs(B.to_host() - Xo).max() <= 200 * EPS * np.abs(Xo).max()


def test_getrs_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "gesv_d.npz"))
    n, nb = 300, 128
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    sl.getrs(A, piv, B)
    X = B.to_host()
    assert np.abs(X - g["out"]).max() <= 1e-10 * np.abs(g["out"]).max()
    assert o.solve_residual(o.generate("rand", n, n, 42), X, o.generate("rand", n, 10, 43)) <= 25 * EPS


@pytest.mark.parametrize("n,nb,nrhs", [(1024, 256, 10), (700, 128, 130), (2048, 512, 3)])
def test_getrs_vs_oracle_and_tester_residual(sl, n, nb, nrhs):
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host()
    sl.getrs(A, piv, B)
    X = B.to_host()
    Xo = o.getrs(LU, piv, o.generate("rand", n, nrhs, 43), nb)
    assert np.abs(X - Xo).max() <= 1e-9 * np.abs(Xo).max()
    assert o.solve_residual(o.generate("rand", n, n, 42), X, o.generate("rand", n, nrhs, 43)) <= 25 * EPS


def test_hemm_z_matches_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "hemm_z.npz"))["out"]
    n, nb, nrhs = 192, 64, 70
    A = sl.HermitianMatrix(n, nb, dtype="z").generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype="z").generate("rand", 43)
    C = sl.Matrix(n, nrhs, nb, dtype="z").generate("rand", 44)
    sl.hemm(ALPHA, A, B, BETA, C)
    assert np.abs(C.to_host() - g).max() <= 8 * np.sqrt(n) * EPS * np.abs(g).max()


@pytest.mark.parametrize("t,n,nb,nrhs,beta", [("d", 1000, 128, 10, 1.0), ("d", 512, 256, 300, 0.0), ("s", 300, 64, 7, -0.5)])
def test_hemm_and_norm_vs_oracle(sl, t, n, nb, nrhs, beta):
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 44)
    sl.hemm(-1.0, A, B, beta, C)
    hi = np.float64
    a = o.generate("rand_dominant", n, n, 42, NP[t]).astype(hi)
    ref = o.hemm(-1.0, a, o.generate("rand", n, nrhs, 43, NP[t]).astype(hi), beta, o.generate("rand", n, nrhs, 44, NP[t]).astype(hi), nb)
    scale = np.abs(o.he_full(a)) @ np.abs(o.generate("rand", n, nrhs, 43, NP[t]).astype(hi))
    assert (np.abs(C.to_host() - ref) <= 4 * np.sqrt(n) * TOL[t] * (scale + 1.0)).all()
    assert abs(sl.norm_inf(A) - o.norm_inf(a, True)) <= 4 * n * TOL[t] * o.norm_inf(a, True)
    G = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    gref = o.norm_inf(o.generate("rand", n, nrhs, 43, NP[t]).astype(hi))
    assert abs(sl.norm_inf(G) - gref) <= 4 * nrhs * TOL[t] * gref


@pytest.mark.parametrize("n,nb,tc", [(1024, 256, False), (1024, 256, True), (1536, 512, True), (700, 128, True), (300, 512, True)])
def test_getrf_float_paths(sl, n, nb, tc):
    """FP32 LU (SIMT and tcgen05 trailing update): ||P A - L U|| / (n ||A||) at FP32 level, and the solve through
    getrs_s passes the tester's residual check at FP32."""
    A = sl.Matrix(n, n, nb, dtype="s").generate("rand", 42)
    piv, info = sl.getrf(A, {"tensor_core_fp32": tc})
    assert info == 0
    LU = A.to_host().astype(np.float64)
    a = o.generate("rand", n, n, 42, np.float32).astype(np.float64)
    perm = o.pivots_to_perm(piv, n, nb)
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    assert np.abs(L).max() <= 1.0 + 1e-6                       # partial pivoting: |l_ij| <= 1
    resid = np.abs(a[perm] - L @ U).max() / (n * np.abs(a).max())
    assert resid <= 3 * EPS32
    B = sl.Matrix(n, 10, nb, dtype="s").generate("rand", 43)
    sl.getrs(A, piv, B)
    b = o.generate("rand", n, 10, 43, np.float32).astype(np.float64)
    X = B.to_host().astype(np.float64)
    r = np.abs(b - a @ X).sum(axis=0).max() / (n * np.abs(a).sum(axis=0).max() * np.abs(X).sum(axis=0).max())
    assert r <= 25 * EPS32                                     # test/test_gesv.cc residual, float tolerance


@pytest.mark.parametrize("tc05", ["1", "0"])
def test_posv_mixed_matches_reference_golden(sl, golden_dir, monkeypatch, tc05):
    monkeypatch.setenv("SB200_MIXED_TC05", tc05)
    g = np.load(os.path.join(golden_dir, "posv_mixed_d.npz"))
    n, nb = 256, 64
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    a_before = A.to_host()
    info, it, tm = sl.posv_mixed(A, B, X)
    assert info == int(g["info"]) == 0
    if tc05 == "0":
        assert it == int(g["iters"])                # FP32 SIMT factor: the reference's iteration count
    else:
        assert 0 <= it <= int(g["iters"]) + 1       # FP32-emulated factor: same accuracy class
    x = X.to_host()
    assert np.abs(x - g["out"]).max() <= 1e-13 * np.abs(g["out"]).max()
    assert np.array_equal(A.to_host(), a_before)    # converged: A is not overwritten
    assert tm["total"] > 0 and tm["factor_lo"] > 0


@pytest.mark.parametrize("tc05", ["1", "0"])
def test_gesv_mixed_matches_reference_golden(sl, golden_dir, monkeypatch, tc05):
    monkeypatch.setenv("SB200_MIXED_TC05", tc05)
    g = np.load(os.path.join(golden_dir, "gesv_mixed_d.npz"))
    n, nb = 256, 64
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    info, it, piv, tm = sl.gesv_mixed(A, B, X)
    assert info == int(g["info"]) == 0
    assert 0 <= it <= int(g["iters"]) + 1
    x = X.to_host()
    assert np.abs(x - g["out"]).max() <= 1e-11 * np.abs(g["out"]).max()
    assert o.solve_residual(o.generate("rand", n, n, 42), x, o.generate("rand", n, 10, 43)) <= 25 * EPS


@pytest.mark.parametrize("routine,n,nb", [("posv", 2048, 512), ("gesv", 2048, 512), ("posv", 1000, 128), ("gesv", 700, 128),
                                          ("posv", 4096, 512), ("gesv", 4096, 512)])
def test_mixed_solvers_vs_oracle_and_tester_residual(sl, routine, n, nb):
    herm = routine == "posv"
    kind = "rand_dominant" if herm else "rand"
    A = (sl.HermitianMatrix(n, nb) if herm else sl.Matrix(n, n, nb)).generate(kind, 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    res = sl.posv_mixed(A, B, X) if herm else sl.gesv_mixed(A, B, X)
    info, it = res[0], res[1]
    assert info == 0 and 0 <= it <= 30
    a = o.generate(kind, n, n, 42); b = o.generate("rand", n, 10, 43)
    af = o.he_full(a) if herm else a
    x = X.to_host()
    assert o.solve_residual(af, x, b) <= 25 * EPS               # test/test_posv.cc:336-342, test_gesv.cc
    if n <= 1000:
        xo, ito, _ = o.solve_mixed(a, b, nb, hermitian=herm)
        assert abs(it - ito) <= 1
        assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()


def test_mixed_failure_codes_and_fallback(sl):
    """iter = -3 when the low-precision factorisation fails (info from the FP64 fallback), -(itermax+1) when the
    refinement does not converge; the FP64 fallback then solves the system (src/posv_mixed.cc:262-287)."""
    n, nb = 512, 128
    rng = np.random.default_rng(3)
    Q = rng.random((n, n))
    S = Q @ Q.T + n * np.eye(n)
    S[300, 300] = -1.0
    A = sl.HermitianMatrix(n, nb); A.from_host(np.asfortranarray(S))
    B = sl.Matrix(n, 3, nb).generate("rand", 1); X = sl.Matrix(n, 3, nb)
    info, it, _ = sl.posv_mixed(A, B, X)
    assert it == -3 and info == 301
    _, ito, infoo = o.solve_mixed(np.tril(S), o.generate("rand", n, 3, 1), nb, hermitian=True)
    assert (ito, infoo) == (it, info)
    A = sl.Matrix(n, n, nb).generate("rand", 1)
    info, it, piv, tm = sl.gesv_mixed(A, B, X, {"max_iterations": 0, "tolerance": 1e-30})
    assert it == -1 and info == 0 and tm["factor_hi"] > 0
    a = o.generate("rand", n, n, 1)
    assert o.solve_residual(a, X.to_host(), o.generate("rand", n, 3, 1)) <= 25 * EPS
    info, it, piv, tm = sl.gesv_mixed(sl.Matrix(n, n, nb).generate("rand", 1), B, X,
                                      {"max_iterations": 0, "tolerance": 1e-30, "use_fallback_solver": False})
    assert it == -1 and tm["factor_hi"] == 0


def test_solve_path_rejects_multi_rank_and_type_mismatch(sl):
    A = sl.HermitianMatrix(64, 32)
    Bs = sl.Matrix(64, 4, 32, dtype="s")
    with pytest.raises(sl.SB200Error):
        sl.potrs(A, Bs)
    with pytest.raises(sl.SB200Error):
        sl.posv_mixed(A, sl.Matrix(64, 4, 16), sl.Matrix(64, 4, 32))


# ---------------------------------------------------------------------------- small-nrhs solve kernel (trsm_small)
@pytest.mark.parametrize("t,n,nb,nrhs", [("d", 1000, 128, 5), ("z", 520, 128, 10), ("s", 1024, 256, 10), ("c", 300, 64, 3),
                                         ("d", 1536, 512, 64), ("d", 200, 512, 9), ("z", 1100, 512, 8)])
def test_potrs_small_rhs_kernel_all_types_and_ragged(sl, t, n, nb, nrhs):
    """nrhs <= 64 takes the one-launch-per-step tile solve (all three sweeps it serves: lower N, lower C, and --
    through getrs below -- upper N), ragged last tiles and every scalar type; vs the oracle's block sweeps."""
    hi = np.complex128 if t in "cz" else np.float64
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host()).astype(hi)
    sl.potrs(A, B)
    Xo = o.potrs(L, o.generate("rand", n, nrhs, 43, NP[t]).astype(hi), nb)
    assert np.abs(B.to_host() - Xo).max() <= 200 * TOL[t] * np.abs(Xo).max()


@pytest.mark.parametrize("t,n,nb,nrhs", [("d", 700, 128, 7), ("s", 1024, 256, 10), ("d", 1536, 512, 33), ("d", 300, 512, 2)])
def test_getrs_small_rhs_kernel(sl, t, n, nb, nrhs):
    A = sl.Matrix(n, n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host().astype(np.float64)
    sl.getrs(A, piv, B)
    Xo = o.getrs(LU, piv, o.generate("rand", n, nrhs, 43, NP[t]).astype(np.float64), nb)
    tol = 1e-9 if t == "d" else 0.5           # forward error ~ cond(A) eps (sanity only in FP32); the tester residual is the sharp check:
    X = B.to_host().astype(np.float64)
    assert np.abs(X - Xo).max() <= tol * np.abs(Xo).max()
    a = o.generate("rand", n, n, 42, NP[t]).astype(np.float64); b = o.generate("rand", n, nrhs, 43, NP[t]).astype(np.float64)
    assert o.solve_residual(a, X, b) <= 25 * TOL[t]


def test_empty_problems_are_quick_returns(sl):
    """n == 0 / nrhs == 0: LAPACK-style quick returns, no launches that index an empty plan."""
    A = sl.HermitianMatrix(0, 64)
    assert sl.potrf(A) == 0
    B0 = sl.Matrix(0, 0, 64)
    sl.potrs(A, B0)
    sl.gemm(1.0, sl.Matrix(0, 0, 64), sl.Matrix(0, 0, 64), 1.0, sl.Matrix(0, 0, 64))
    n, nb = 256, 64
    H = sl.HermitianMatrix(n, nb).generate("rand_dominant", 1)
    assert sl.potrf(H) == 0
    Bn = sl.Matrix(n, 0, nb)
    sl.potrs(H, Bn)
    G = sl.Matrix(n, n, nb).generate("rand", 1)
    piv, info = sl.getrf(G)
    sl.getrs(G, piv, Bn)
    assert sl.Matrix(0, 0, 64).to_host().shape == (0, 0)
