"""CPU tests: the numpy oracle (oracle/slate_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/*.npz <- oracle/_ref/ref_dump, see tests/golden/make_golden.py).
This is what "pins" the oracle: Philox and pivots bit-exact, floating point to a few ulp."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o

EPS = np.finfo(np.float64).eps
ALPHA = 3.141592653589793 + 1.414213562373095j     # tester defaults (test/test.cc:447-448)
BETA = 2.718281828459045 + 1.732050807568877j


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("t,dtype", [("s", np.float32), ("d", np.float64), ("c", np.complex64), ("z", np.complex128)])
def test_philox_bit_exact(golden_dir, t, dtype):
    g = load(golden_dir, f"gen_{t}")
    assert np.array_equal(o.generate("rand", 80, 96, 5, dtype), g["rand"])
    assert np.array_equal(o.generate("rand_dominant", 96, 96, 7, dtype), g["rand_dominant"])


def test_philox_offset_blocks_are_distribution_independent():
    full = o.generate("rand", 70, 50, 9)
    blk = o.generate("rand", 20, 10, 9, i0=30, j0=25)
    assert np.array_equal(full[30:50, 25:35], blk)


def test_potrf_matches_reference(golden_dir):
    g = load(golden_dir, "potrf_d")
    n, nb = 384, 128
    G = o.generate("rand_dominant", n, n, 42)
    A = np.tril(G) + np.tril(G, -1).T
    L, info = o.potrf(A, nb)
    assert info == int(g["info"]) == 0
    ref = np.tril(g["out"])
    assert np.abs(L - ref).max() <= 16 * EPS * np.abs(ref).max()


def test_potrf_info_not_positive_definite():
    n, nb = 96, 32
    A = np.eye(n); A[50, 50] = -1.0
    _, info = o.potrf(A, nb)
    assert info == 51


@pytest.mark.parametrize("name,n", [("getrf_d", 384), ("getrf_d_ragged", 300)])
def test_getrf_matches_reference_with_identical_pivots(golden_dir, name, n):
    g = load(golden_dir, name)
    A = o.generate("rand", n, n, 42)
    LU, piv, info = o.getrf(A, 128, 16)
    flat = np.array([p for col in piv for p in col], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference"
    assert info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()
    # P A = L U
    perm = o.pivots_to_perm(piv, n, 128)
    L = np.tril(LU, -1) + np.eye(n)
    U = np.triu(LU)
    assert np.abs(A[perm] - L @ U).max() <= 1e-13 * n


def test_getrf_zero_pivot_sets_info():
    A = o.generate("rand", 64, 64, 3)
    A[:, 10] = 0.0
    _, _, info = o.getrf(A, 32, 8)
    assert info == 11


@pytest.mark.parametrize("t,dtype", [("d", np.float64), ("z", np.complex128)])
def test_gemm_matches_reference(golden_dir, t, dtype):
    g = load(golden_dir, f"gemm_{t}")
    n, nb = 256, 64
    A, B, C = (o.generate("rand", n, n, s, dtype) for s in (42, 43, 44))
    al, be = (ALPHA, BETA) if t == "z" else (ALPHA.real, BETA.real)
    out = o.gemm(al, A, B, be, C, nb)
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    assert o.gemm_check(al, A, B, be, C, g["out"]) <= 3 * EPS      # reference output passes its own test


@pytest.mark.parametrize("t,dtype", [("d", np.float64), ("z", np.complex128)])
def test_herk_matches_reference(golden_dir, t, dtype):
    g = load(golden_dir, f"herk_{t}")
    n, k, nb = 256, 128, 64
    A = o.generate("rand", n, k, 42, dtype)
    C = np.tril(o.generate("rand", n, n, 44, dtype))
    out = np.tril(o.herk(ALPHA.real, A, BETA.real, C, nb))
    ref = np.tril(g["out"])
    assert np.abs(out - ref).max() <= 64 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("t,dtype", [("d", np.float64), ("z", np.complex128)])
def test_her2k_matches_reference(golden_dir, t, dtype):
    g = load(golden_dir, f"her2k_{t}")
    n, k, nb = 200, 100, 64
    A = o.generate("rand", n, k, 42, dtype)
    B = o.generate("rand", n, k, 43, dtype)
    C = np.tril(o.generate("rand", n, n, 44, dtype))
    al = ALPHA if t == "z" else ALPHA.real
    out = np.tril(o.her2k(al, A, B, BETA.real, C, nb))
    ref = np.tril(g["out"])
    assert np.abs(out - ref).max() <= 64 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("routine", ["syrk", "syr2k"])
def test_complex_symmetric_rank_updates_match_reference(golden_dir, routine):
    g = load(golden_dir, f"{routine}_z")
    n, k, nb = 200, 100, 64
    A = o.generate("rand", n, k, 42, np.complex128)
    B = o.generate("rand", n, k, 43, np.complex128)
    C = np.tril(o.generate("rand", n, n, 44, np.complex128))
    out = o.syrk(ALPHA, A, BETA, C, nb) if routine == "syrk" else o.syr2k(ALPHA, A, B, BETA, C, nb)
    ref = np.tril(g["out"])
    assert np.abs(np.tril(out) - ref).max() <= 64 * EPS * np.abs(ref).max()


def test_trmm_matches_reference(golden_dir):
    g = load(golden_dir, "trmm_d")
    m, n, nb = 200, 70, 64
    A = np.tril(o.generate("rand", m, m, 42))
    B = o.generate("rand", m, n, 43)
    out = o.trmm(ALPHA.real, A, B, nb)
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("name,t,side,op,unit", [
    ("trmm_z_left_conj", "z", "L", "C", False), ("trmm_d_left_trans", "d", "L", "T", True),
    ("trmm_d_right", "d", "R", "N", False), ("trmm_z_right_trans", "z", "R", "T", False),
    ("trmm_z_right_conj", "z", "R", "C", True)])
def test_trmm_side_op_variants_match_reference(golden_dir, name, t, side, op, unit):
    """slate::trmm with Side::Right and with (conjugate-)transposed views of a lower-triangular A: the unmodified
    reference's outputs (tests/golden/make_golden.py blas3_variants)."""
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    (m, n), nb = ((200, 70) if side == "L" else (70, 200)), 64
    na = m if side == "L" else n
    A = np.tril(o.generate("rand", na, na, 42, dt))
    B = o.generate("rand", m, n, 43, dt)
    out = o.trmm(ALPHA if t == "z" else ALPHA.real, A, B, nb, unit=unit, side=side, op=op)
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("name,t,n,op", [("gesv_d_trans", "d", 300, "T"), ("gesv_z_conj", "z", 200, "C")])
def test_getrs_with_transposed_view_matches_reference(golden_dir, name, t, n, op):
    """lu_solve_using_factor handed a (conjugate-)transposed view of the factored matrix: op(A) X = B (src/getrs.cc:97-112)."""
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    nb, nrhs = 64, 70
    A = o.generate("rand", n, n, 42, dt)
    B = o.generate("rand", n, nrhs, 43, dt)
    LU, piv, info = o.getrf(A, nb)
    assert info == 0
    X = o.getrs(LU, piv, B, nb, op=op)
    assert np.abs(X - g["out"]).max() <= 1e-10 * np.abs(g["out"]).max()
    M = A.T if op == "T" else A.conj().T
    assert o.solve_residual(M, X, B) <= 25 * EPS


@pytest.mark.parametrize("name,routine,t", [("herk_z_conj", "herk", "z"), ("herk_d_trans", "herk", "d"), ("her2k_z_conj", "her2k", "z"),
                                            ("syrk_z_trans", "syrk", "z"), ("syr2k_z_trans", "syr2k", "z")])
def test_rank_updates_with_transposed_views_match_reference(golden_dir, name, routine, t):
    """slate::herk / her2k / syrk / syr2k handed (conjugate-)transposed views of A (and B) stored k x n: the same update as
    with the n x k matrix A^H (A^T), so the restatements are called on that."""
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    n, k, nb = 200, 100, 64
    A0 = o.generate("rand", k, n, 42, dt)
    B0 = o.generate("rand", k, n, 43, dt)
    C = np.tril(o.generate("rand", n, n, 44, dt))
    if routine in ("herk", "her2k"):
        A, B = A0.conj().T, B0.conj().T
    else:
        A, B = A0.T, B0.T
    if routine == "herk":
        out = o.herk(ALPHA.real, A, BETA.real, C, nb)
    elif routine == "her2k":
        out = o.her2k(ALPHA if t == "z" else ALPHA.real, A, B, BETA.real, C, nb)
    elif routine == "syrk":
        out = o.syrk(ALPHA, A, BETA, C, nb)
    else:
        out = o.syr2k(ALPHA, A, B, BETA, C, nb)
    ref = np.tril(g["out"])
    assert np.abs(np.tril(out) - ref).max() <= 64 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("name,t,opa,opb", [("gemm_d_tn", "d", "T", "N"), ("gemm_d_nt", "d", "N", "T"), ("gemm_z_cn", "z", "C", "N"),
                                           ("gemm_z_tc", "z", "T", "C"), ("gemm_z_nc", "z", "N", "C")])
def test_gemm_with_transposed_views_matches_reference(golden_dir, name, t, opa, opb):
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    m, n, k, nb = 150, 200, 100, 64
    A = o.generate("rand", *((m, k) if opa == "N" else (k, m)), 42, dt)
    B = o.generate("rand", *((k, n) if opb == "N" else (n, k)), 43, dt)
    C = o.generate("rand", m, n, 44, dt)
    al, be = (ALPHA, BETA) if t == "z" else (ALPHA.real, BETA.real)
    out = o.gemm(al, A, B, be, C, nb, opa=opa, opb=opb)
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("name,t,side,op,unit", [
    ("trsm_z_left_conj", "z", "L", "C", False), ("trsm_d_left_trans", "d", "L", "T", True),
    ("trsm_d_right", "d", "R", "N", False), ("trsm_z_right_trans", "z", "R", "T", True),
    ("trsm_z_right_conj", "z", "R", "C", False)])
def test_trsm_side_op_variants_match_reference(golden_dir, name, t, side, op, unit):
    """slate::trsm with Side::Right and with (conjugate-)transposed views of a lower-triangular A (rand_dominant)."""
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    (m, n), nb = ((200, 70) if side == "L" else (70, 200)), 64
    na = m if side == "L" else n
    A = np.tril(o.generate("rand_dominant", na, na, 42, dt))
    B = o.generate("rand", m, n, 43, dt)
    out = o.trsm(ALPHA if t == "z" else ALPHA.real, A, B, nb, side=side, lower=True, op=op, unit=unit)
    assert out.shape == g["out"].shape
    # a unit-diagonal solve with a rand lower triangle grows: compare relative to the largest entry
    assert np.abs(out - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()


@pytest.mark.parametrize("name,routine,t,n", [("hemm_z_right", "hemm", "z", 192), ("hemm_d_right", "hemm", "d", 200),
                                              ("symm_z_right", "symm", "z", 192)])
def test_hemm_symm_right_match_reference(golden_dir, name, routine, t, n):
    g = load(golden_dir, name)
    dt = np.complex128 if t == "z" else np.float64
    nb, nrhs = 64, 70
    A = np.tril(o.generate("rand", n, n, 42, dt))
    B = o.generate("rand", nrhs, n, 43, dt)
    C = o.generate("rand", nrhs, n, 44, dt)
    al, be = (ALPHA, BETA) if t == "z" else (ALPHA.real, BETA.real)
    out = (o.hemm if routine == "hemm" else o.symm)(al, A, B, be, C, nb, side="R")
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


def test_symm_matches_reference(golden_dir):
    g = load(golden_dir, "symm_z")
    n, nb, nrhs = 192, 64, 70
    A = np.tril(o.generate("rand", n, n, 42, np.complex128))
    B = o.generate("rand", n, nrhs, 43, np.complex128)
    C = o.generate("rand", n, nrhs, 44, np.complex128)
    out = o.symm(ALPHA, A, B, BETA, C, nb)
    assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


def test_getrf_nopiv_matches_reference(golden_dir):
    g = load(golden_dir, "getrf_nopiv_d")
    n, nb = 300, 128
    A = o.generate("rand_dominant", n, n, 42)
    LU, info = o.getrf_nopiv(A, nb)
    assert info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    assert np.abs(L @ U - A).max() <= 64 * EPS * np.abs(A).max()
    Z = A.copy(); Z[:, 5] = 0.0; Z[5, 5] = 0.0; Z[5:, 5] = 0.0            # exact zero pivot at column 5 after elimination
    Z[:5, 5] = 0.0
    assert o.getrf_nopiv(Z, nb)[1] == 6


@pytest.mark.parametrize("name,m,n", [("getrf_tntpiv_d", 384, 384), ("getrf_tntpiv_d_ragged", 300, 300),
                                      ("getrf_tntpiv_d_tall", 512, 256)])
def test_getrf_tntpiv_matches_reference(golden_dir, name, m, n):
    """MethodLU::CALU on one rank (src/getrf_tntpiv.cc): the pivots of partial pivoting, L21 through a solve with U11."""
    g = load(golden_dir, name)
    A = o.generate("rand", m, n, 42)
    LU, piv, info = o.getrf_tntpiv(A, 128, 16)
    flat = np.array([p for col in piv for p in col], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference"
    assert info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()


@pytest.mark.parametrize("ranks", [2, 3, 4])
@pytest.mark.parametrize("m,n,nb", [(384, 384, 64), (300, 300, 64), (448, 256, 64)])
def test_getrf_tntpiv_tournament_is_an_lu(ranks, m, n, nb):
    """Several ranks per panel (restated by reading; the reference cannot run multi-rank here): P A = L U holds, the
    winners of every panel are rows of that panel, and the growth |L| stays small."""
    A = o.generate("rand", m, n, 42)
    LU, piv, info = o.getrf_tntpiv(A, nb, 16, ranks=ranks)
    assert info == 0
    mn = min(m, n)
    perm = o.pivots_to_perm(piv, m, nb)
    assert sorted(perm) == list(range(m))
    L = np.tril(LU, -1)[:, :mn] + np.eye(m, mn)
    U = np.triu(LU)[:mn]
    assert np.abs(A[perm] - L @ U).max() <= 1e-13 * m
    assert np.abs(L).max() < 8.0


def test_tnt_winners_to_sequential_is_the_reference_example():
    # internal_getrf_tntpiv.cc:69-93: mt = 2, nb = 4, winners (1,1) (0,3) (0,0) (0,2) -> pivots (1,1) (0,3) (1,1) (1,1)
    piv, row_at = o.tnt_winners_to_sequential([5, 3, 0, 2], 8)
    assert list(piv) == [5, 3, 5, 5]
    assert list(row_at[:4]) == [5, 3, 0, 2]


@pytest.mark.parametrize("name,n,dt", [("getrf_z", 192, np.complex128), ("getrf_c", 200, np.complex64)])
def test_complex_getrf_matches_reference_with_identical_pivots(golden_dir, name, n, dt):
    """cabs1 pivot rule (|re| + |im|, src/internal/Tile_getrf.hh:210-237), complex reciprocal scaling."""
    g = load(golden_dir, name)
    A = o.generate("rand", n, n, 42, dtype=dt)
    LU, piv, info = o.getrf(A, 64, 16)
    flat = np.array([p for col in piv for p in col], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "pivot vectors differ from the reference"
    assert info == int(g["info"]) == 0
    eps = np.finfo(dt).eps
    assert np.abs(LU - g["out"]).max() <= 1024 * eps * np.abs(g["out"]).max()


def test_complex_gesv_matches_reference(golden_dir):
    g = load(golden_dir, "gesv_z")
    n, nb, nrhs = 200, 64, 70
    A = o.generate("rand", n, n, 42, dtype=np.complex128)
    B = o.generate("rand", n, nrhs, 43, dtype=np.complex128)
    LU, piv, info = o.getrf(A, nb, 16)
    X = o.getrs(LU, piv, B, nb)
    assert info == int(g["info"]) == 0
    assert np.abs(X - g["out"]).max() <= 1e-11 * np.abs(g["out"]).max()


def test_complex_tntpiv_and_nopiv_match_reference(golden_dir):
    g = load(golden_dir, "getrf_tntpiv_z")
    A = o.generate("rand", 192, 192, 42, dtype=np.complex128)
    LU, piv, info = o.getrf_tntpiv(A, 64, 16)
    assert np.array_equal(np.array([p for col in piv for p in col], dtype=np.int64), g["piv"]) and info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()
    g = load(golden_dir, "getrf_nopiv_z")
    A = o.generate("rand_dominant", 200, 200, 42, dtype=np.complex128)
    LU, info = o.getrf_nopiv(A, 64)
    assert info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


# ----------------------------------------------------------------------------------------------------------------
# The reference ON PROCESS GRIDS: tests/golden/grid_*.npz were written by the unmodified reference running on p x q ranks
# (oracle/_ref/ref_dump_mp = the reference built against the multi-process MPI replacement oracle/mpi_mp, started by
# oracle/mprun.py).  They pin what one rank cannot: the tournament proper (>= 2 ranks in a process column), the cross-rank
# pivot rule of getrf, the grid data flow of potrf.
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,ranks,m,n", [("grid_getrf_tntpiv_d_2x1", 2, 384, 384), ("grid_getrf_tntpiv_d_3x1_ragged", 3, 300, 300),
                                            ("grid_getrf_tntpiv_d_4x1_tall", 4, 448, 256), ("grid_getrf_tntpiv_d_2x4", 2, 512, 512)])
def test_tournament_matches_the_multirank_reference(golden_dir, name, ranks, m, n):
    """internal_getrf_tntpiv.cc with several participants per panel: IDENTICAL pivots, factor to rounding."""
    g = load(golden_dir, name)
    A = o.generate("rand", m, n, 42)
    LU, piv, info = o.getrf_tntpiv(A, 64, 16, ranks=ranks)
    flat = np.array([p for col in piv for p in col], dtype=np.int64)
    assert np.array_equal(flat, g["piv"]), "the restated tournament picks other rows than the reference on this grid"
    assert info == int(g["info"]) == 0
    if "out" in g:
        assert np.abs(LU - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()
    else:
        assert np.abs(np.diag(LU) - g["diag_u"]).max() <= 1e-12 * np.abs(g["diag_u"]).max()
        assert abs(np.abs(LU).sum() - float(g["abs_sum"])) <= 1e-11 * float(g["abs_sum"])


def test_getrf_and_potrf_match_the_multirank_reference(golden_dir):
    """Partial pivoting across ranks (MPI_MAXLOC keeps the lowest rank on ties, Tile_getrf.hh:256-289) and Cholesky on a
    2 x 2 grid give what the one-rank run gives: the restatement needs no grid parameter."""
    g = load(golden_dir, "grid_getrf_d_2x2")
    A = o.generate("rand", 384, 384, 42)
    LU, piv, info = o.getrf(A, 64, 16)
    assert np.array_equal(np.array([p for col in piv for p in col], dtype=np.int64), g["piv"]) and info == int(g["info"]) == 0
    assert np.abs(LU - g["out"]).max() <= 1e-12 * np.abs(g["out"]).max()
    g = load(golden_dir, "grid_getrf_d_3x2_tall")
    LU, piv, info = o.getrf(o.generate("rand", 500, 300, 7), 64, 16)
    assert np.array_equal(np.array([p for col in piv for p in col], dtype=np.int64), g["piv"])
    assert np.abs(np.diag(LU) - g["diag_u"]).max() <= 1e-12 * np.abs(g["diag_u"]).max()
    g = load(golden_dir, "grid_potrf_d_2x2")
    G = o.generate("rand_dominant", 384, 384, 42)
    L, info = o.potrf(o.he_full(np.tril(G)), 64)
    assert info == int(g["info"]) == 0
    assert np.abs(np.tril(L) - g["out"]).max() <= 16 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("t,p", [("d", 2), ("z", 2), ("d", 3), ("s", 2)])
def test_live_multirank_reference_agrees_when_built(tmp_path, t, p):
    """When oracle/_ref/ref_dump_mp is present: the reference on p x 1 ranks, live, another size and seed, CALU."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "ref_dump_mp")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_dump_mp not built in this environment")
    import subprocess
    import sys
    n, nb = 200, 32
    dt = {"d": np.float64, "z": np.complex128, "s": np.float32}[t]
    prefix = str(tmp_path / "x")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    subprocess.run([sys.executable, os.path.join(root, "oracle", "mprun.py"), "-n", str(p), "--timeout", "120", exe, "getrf", t,
                    str(n), str(nb), "11", "0", "0", prefix, f"p={p}", "q=1", "ib=16", "pt=1", "method=calu"],
                   check=True, env=env, capture_output=True)
    piv = np.fromfile(prefix + ".r0.piv.bin", dtype=np.int64).reshape(-1, 2)
    parts = [np.fromfile(f"{prefix}.r{r}.out.bin", dtype=dt).reshape(n, n, order="F") for r in range(p)]
    LU, pv, info = o.getrf_tntpiv(o.generate("rand", n, n, 11, dtype=dt), nb, 16, ranks=p)
    if t != "s":                        # float: a near-tie may resolve differently; the factor identity below still holds
        assert np.array_equal(np.array([x for col in pv for x in col], dtype=np.int64), piv)
        for i in range(-(-n // nb)):                                   # tile row i lives on rank i % p
            blk = slice(i * nb, min((i + 1) * nb, n))
            assert np.abs(parts[i % p][blk] - LU[blk]).max() <= 1e-12 * np.abs(LU).max()
    else:
        ref = np.zeros((n, n), dtype=np.float64)
        for i in range(-(-n // nb)):
            blk = slice(i * nb, min((i + 1) * nb, n))
            ref[blk] = parts[i % p][blk]
        pivs = [[tuple(x) for x in piv[k0:k0 + nb]] for k0 in range(0, n, nb)]
        perm = o.pivots_to_perm(pivs, n, nb)
        L = np.tril(ref, -1) + np.eye(n); U = np.triu(ref)
        assert np.abs(o.generate("rand", n, n, 11, dtype=dt).astype(np.float64)[perm] - L @ U).max() <= 64 * np.finfo(np.float32).eps * n


def test_trsm_matches_reference(golden_dir):
    g = load(golden_dir, "trsm_d")
    m, n = 256, 128
    T = np.tril(o.generate("rand_dominant", m, m, 42))
    B = o.generate("rand", m, n, 43)
    X = o.trsm_tile("L", "L", "N", "N", ALPHA.real, T, B)
    assert np.abs(X - g["out"]).max() <= 256 * EPS * np.abs(g["out"]).max()


def test_norms_match_reference(golden_dir):
    g = load(golden_dir, "norms_d")["out"]
    A = o.generate("rand", 200, 136, 42)
    nb = 64
    tiles = [A[i:i + nb, j:j + nb] for j in range(0, 136, nb) for i in range(0, 200, nb)]
    assert o.combine_norm("M", [o.genorm("M", t) for t in tiles]) == g[0]
    one = np.zeros(136); inf = np.zeros(200)
    for j in range(0, 136, nb):
        for i in range(0, 200, nb):
            one[j:j + nb] += o.genorm("O", A[i:i + nb, j:j + nb])
            inf[i:i + nb] += o.genorm("I", A[i:i + nb, j:j + nb])
    assert abs(one.max() - g[1]) <= 8 * EPS * g[1]
    assert abs(inf.max() - g[2]) <= 8 * EPS * g[2]
    fro = o.combine_norm("F", [o.genorm("F", t) for t in tiles])
    assert abs(fro - g[3]) <= 64 * EPS * g[3]


def test_tile_kernel_restatements():
    rng = np.random.default_rng(0)
    A = rng.random((7, 5)); B = rng.random((7, 5))
    assert np.allclose(o.geadd(2.0, A, -1.0, B), 2 * A - B)
    assert np.allclose(o.gescale(3.0, 2.0, A), 1.5 * A)
    S = o.geset(1.0, 9.0, 4, 6)
    assert S[2, 2] == 9.0 and S[1, 3] == 1.0
    Z = o.tzset("L", 0.0, 1.0, A)
    assert Z[3, 1] == 0.0 and Z[1, 1] == 1.0 and Z[0, 4] == A[0, 4]
    assert np.isnan(o.genorm("M", np.array([[1.0, np.nan]])))
    sc, sq = o.genorm("F", A)
    assert np.isclose(sc * np.sqrt(sq), np.linalg.norm(A))
    H = rng.random((6, 6)); Hs = np.tril(H) + np.tril(H, -1).T
    assert np.isclose(o.henorm("M", "L", H), np.abs(Hs).max())
    assert np.allclose(o.henorm("O", "L", H), np.abs(Hs).sum(axis=0))


def test_gesv_mixed_golden_solves_system(golden_dir):
    """The reference's mixed-precision solve passes the tester residual; keeps the fixture honest."""
    g = load(golden_dir, "gesv_mixed_d")
    n = 256
    A = o.generate("rand", n, n, 42); B = o.generate("rand", n, 10, 43)
    assert int(g["info"]) == 0 and int(g["iters"]) >= 0
    assert o.solve_residual(A, g["out"], B) <= 25 * EPS


def test_live_reference_agrees_when_built(ref_dump, tmp_path):
    """When oracle/_ref is present, re-run the reference live at another size/seed."""
    if ref_dump is None:
        pytest.skip("oracle/_ref not built in this environment")
    import subprocess
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    prefix = str(tmp_path / "p")
    subprocess.run([ref_dump, "potrf", "d", "200", "64", "11", "0", "0", prefix], check=True, env=env,
                   capture_output=True)
    ref = np.fromfile(prefix + ".out.bin").reshape(200, 200, order="F")
    G = o.generate("rand_dominant", 200, 200, 11)
    L, info = o.potrf(np.tril(G) + np.tril(G, -1).T, 64)
    assert info == 0 and np.abs(L - np.tril(ref)).max() <= 16 * EPS * np.abs(ref).max()


# ---------------------------------------------------------------------------- solve path / mixed solvers
def test_posv_matches_reference(golden_dir):
    g = load(golden_dir, "posv_d")
    n, nb = 300, 128
    G = o.generate("rand_dominant", n, n, 42); B = o.generate("rand", n, 10, 43)
    L, info = o.potrf(o.he_full(G), nb)
    X = o.potrs(L, B, nb)
    assert info == int(g["info"]) == 0
    assert np.abs(X - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    assert o.solve_residual(o.he_full(G), X, B) <= 25 * EPS


def test_posv_complex_matches_reference(golden_dir):
    g = load(golden_dir, "posv_z")
    n, nb = 192, 64
    G = o.generate("rand_dominant", n, n, 42, np.complex128); B = o.generate("rand", n, 70, 43, np.complex128)
    L, info = o.potrf(o.he_full(G), nb)
    gz = load(golden_dir, "potrf_z")
    assert info == int(gz["info"]) == 0
    assert np.abs(L - np.tril(gz["out"])).max() <= 64 * EPS * np.abs(gz["out"]).max()
    X = o.potrs(L, B, nb)
    assert np.abs(X - g["out"]).max() <= 256 * EPS * np.abs(g["out"]).max()


def test_gesv_matches_reference(golden_dir):
    g = load(golden_dir, "gesv_d")
    n, nb = 300, 128
    A = o.generate("rand", n, n, 42); B = o.generate("rand", n, 10, 43)
    LU, piv, info = o.getrf(A, nb, 16)
    X = o.getrs(LU, piv, B, nb)
    assert info == int(g["info"]) == 0
    assert np.abs(X - g["out"]).max() <= 1e-10 * np.abs(g["out"]).max()
    assert o.solve_residual(A, X, B) <= 25 * EPS


def test_hemm_matches_reference(golden_dir):
    g = load(golden_dir, "hemm_z")
    n, nb, nrhs = 192, 64, 70
    A = o.generate("rand", n, n, 42, np.complex128)
    B = o.generate("rand", n, nrhs, 43, np.complex128); C = o.generate("rand", n, nrhs, 44, np.complex128)
    out = o.hemm(ALPHA, A, B, BETA, C, nb)
    assert np.abs(out - g["out"]).max() <= 8 * np.sqrt(n) * EPS * np.abs(g["out"]).max()


def test_posv_mixed_matches_reference(golden_dir):
    """Same control flow as src/posv_mixed.cc: same iteration count on the seeded system, same solution to FP64
    refinement accuracy."""
    g = load(golden_dir, "posv_mixed_d")
    n, nb = 256, 64
    G = o.generate("rand_dominant", n, n, 42); B = o.generate("rand", n, 10, 43)
    X, it, info = o.solve_mixed(G, B, nb, hermitian=True)
    assert info == int(g["info"]) == 0
    assert it == int(g["iters"])
    assert np.abs(X - g["out"]).max() <= 1e-13 * np.abs(g["out"]).max()
    assert o.solve_residual(o.he_full(G), X, B) <= 25 * EPS


def test_gesv_mixed_matches_reference(golden_dir):
    g = load(golden_dir, "gesv_mixed_d")
    n, nb = 256, 64
    A = o.generate("rand", n, n, 42); B = o.generate("rand", n, 10, 43)
    X, it, info = o.solve_mixed(A, B, nb, hermitian=False)
    assert info == int(g["info"]) == 0
    assert it == int(g["iters"])
    assert np.abs(X - g["out"]).max() <= 1e-11 * np.abs(g["out"]).max()


@pytest.mark.parametrize("routine,kind,herm", [("gesv_mixed", "rand", False), ("posv_mixed", "rand_dominant", True)])
def test_complex_mixed_solvers_match_reference(golden_dir, routine, kind, herm):
    """<complex<double>, complex<float>> instantiations (src/gesv_mixed.cc:303-316): same iteration count, same solution."""
    g = load(golden_dir, routine + "_z")
    n, nb = 256, 64
    A = o.generate(kind, n, n, 42, dtype=np.complex128); B = o.generate("rand", n, 10, 43, dtype=np.complex128)
    X, it, info = o.solve_mixed(np.tril(A) if herm else A, B, nb, hermitian=herm)
    assert info == int(g["info"]) == 0
    assert it == int(g["iters"])
    assert np.abs(X - g["out"]).max() <= 1e-11 * np.abs(g["out"]).max()
    assert o.solve_residual(o.he_full(np.tril(A)) if herm else A, X, B) <= 25 * EPS


def test_mixed_fallback_and_failure_codes():
    """iter = -3 when the low-precision factorisation fails, -(itermax+1) when refinement does not converge."""
    n, nb = 96, 32
    rng = np.random.default_rng(3)
    Q = rng.random((n, n))
    S = Q @ Q.T + n * np.eye(n)
    S[5, 5] = -1.0                                           # not positive definite: potrf info = 6
    X, it, info = o.solve_mixed(np.tril(S), rng.random((n, 3)), nb, hermitian=True, use_fallback=False)
    assert it == -3 and info == 6
    A = o.generate("rand", n, n, 1)
    X, it, info = o.solve_mixed(A, rng.random((n, 3)), nb, hermitian=False, itermax=0, tol=1e-30, use_fallback=True)
    assert it == -1 and info == 0
    assert o.solve_residual(A, X, (A @ X)) <= 25 * EPS       # fallback solved in FP64


# one-rank golden file, routine, type, n, nb, extra keys, rows of the output, tolerance (relative to the largest entry)
_GRID_CASES = [
    ("potrf_d", "potrf", "d", 384, 128, {}, 384, 64 * EPS),
    ("getrf_d_ragged", "getrf", "d", 300, 128, {"ib": 16, "pt": 1}, 300, 1e-12),
    ("getrf_z", "getrf", "z", 192, 64, {"ib": 16, "pt": 1}, 192, 1e-12),
    ("gemm_d", "gemm", "d", 256, 64, {}, 256, 64 * EPS),
    ("herk_z", "herk", "z", 256, 64, {"k": 128}, 256, 64 * EPS),
    ("trsm_d", "trsm", "d", 128, 64, {"m": 256}, 256, 1e-12),
    ("posv_d", "posv", "d", 300, 128, {}, 300, 1e-12),
    ("gesv_d", "gesv", "d", 300, 128, {"ib": 16, "pt": 1}, 300, 1e-11),
    ("hemm_z", "hemm", "z", 192, 64, {"nrhs": 70}, 192, 64 * EPS),
    ("her2k_z", "her2k", "z", 200, 64, {"k": 100}, 200, 64 * EPS),
    ("getrf_nopiv_d", "getrf_nopiv", "d", 300, 128, {}, 300, 1e-12),
    ("gesv_mixed_d", "gesv_mixed", "d", 256, 64, {"ib": 16, "pt": 1}, 256, 1e-11),
    ("posv_mixed_z", "posv_mixed", "z", 256, 64, {}, 256, 1e-11),
]


@pytest.mark.parametrize("golden,routine,t,n,nb,kv,rows,tol", _GRID_CASES, ids=[c[0] for c in _GRID_CASES])
def test_reference_on_a_2x2_grid_reproduces_its_one_rank_golden(golden_dir, golden, routine, t, n, nb, kv, rows, tol):
    """Live, when oracle/_ref/ref_dump_mp is present: the unmodified reference on a 2 x 2 process grid (oracle/mpi_mp)
    gives, tile by tile, what it gave on one rank (the committed fixture) -- same pivots, same refinement iteration counts,
    outputs to rounding.  This is what licenses comparing the multi-GPU runs of the product with the one-rank oracle."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "ref_dump_mp")):
        pytest.skip("oracle/_ref/ref_dump_mp not built in this environment")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = load(golden_dir, golden)
    kv = dict(kv)
    m = kv.pop("m", None)
    f, meta = mg.run_mp(routine, t, n, nb, 2, 2, m=m, rows=rows, **kv)
    out = f["out"]
    ref = g["out"]
    if routine in ("potrf", "herk", "her2k"):
        out, ref = np.tril(out), np.tril(ref)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()
    if "piv" in g:
        assert np.array_equal(f["piv"].reshape(-1, 2), g["piv"])
    if "iters" in g:
        assert int(meta["iters"]) == int(g["iters"])
    if "info" in g:
        assert int(meta["info"]) == int(g["info"])


@pytest.mark.parametrize("grid,routine,extra", [("2x2", "potrf", []), ("2x2", "getrf", ["--method-lu", "PPLU,CALU"]), ("2x4", "getrf", ["--method-lu", "CALU"]),
                                                ("2x2", "gemm", []), ("2x2", "gesv_mixed", []), ("3x1", "getrf", ["--method-lu", "CALU", "--type", "z"])])
def test_reference_tester_passes_on_process_grids(grid, routine, extra):
    """The reference's OWN acceptance checks (test/test_posv.cc:304-345, test_gesv.cc:371-377, test_gemm.cc:205-207) pass when the
    unmodified reference runs on p x q ranks over the multi-process MPI replacement (oracle/_ref/tester_mp): the replacement
    is faithful enough to carry the grid golden vectors."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "tester_mp")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/tester_mp not built in this environment")
    import subprocess
    import sys
    p, q = (int(x) for x in grid.split("x"))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, os.path.join(root, "oracle", "mprun.py"), "-n", str(p * q), "--timeout", "150", exe, "--grid", grid,
           "--type", "d", "--dim", "300", "--nb", "64", "--target", "t", "--check", "y", "--ref", "n"] + extra + [routine]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=200)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-400:]
    assert f"All tests passed: {routine}" in r.stdout
    assert "FAILED" not in r.stdout


def _sweep_cases():
    rng = np.random.default_rng(20261018)
    out = []
    for _ in range(10):
        nb = int(rng.choice([16, 24, 32, 48]))
        p, q = int(rng.integers(2, 5)), int(rng.integers(1, 3))
        if rng.random() < 0.5:
            n = int(rng.integers(3, 9)) * nb + int(rng.integers(0, nb))          # square, ragged last tile
            m = n
        else:
            n = int(rng.integers(2, 6)) * nb                                      # tall, full-width panels
            m = n + int(rng.integers(1, 4 * nb))
        out.append((m, n, nb, p, q, int(rng.integers(1, 1000))))
    return out


@pytest.mark.parametrize("m,n,nb,p,q,seed", _sweep_cases())
def test_tournament_sweep_against_the_live_multirank_reference(tmp_path, m, n, nb, p, q, seed):
    """Seeded sweep over shapes, tile sizes and grids: the restated tournament picks the reference's pivots everywhere."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "ref_dump_mp")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_dump_mp not built in this environment")
    import subprocess
    import sys
    prefix = str(tmp_path / "x")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    subprocess.run([sys.executable, os.path.join(root, "oracle", "mprun.py"), "-n", str(p * q), "--timeout", "150", exe, "getrf", "d",
                    str(n), str(nb), str(seed), "0", "0", prefix, f"p={p}", f"q={q}", f"m={m}", "ib=8", "pt=1", "method=calu"],
                   check=True, env=env, capture_output=True, timeout=200)
    piv = np.fromfile(prefix + ".r0.piv.bin", dtype=np.int64).reshape(-1, 2)
    LU, pv, info = o.getrf_tntpiv(o.generate("rand", m, n, seed), nb, 8, ranks=p)
    assert np.array_equal(np.array([x for col in pv for x in col], dtype=np.int64), piv)
    parts = [np.fromfile(f"{prefix}.r{r}.out.bin").reshape(m, n, order="F") for r in range(p * q)]
    for j in range(-(-n // nb)):
        for i in range(-(-m // nb)):
            ri, cj = slice(i * nb, min((i + 1) * nb, m)), slice(j * nb, min((j + 1) * nb, n))
            assert np.abs(parts[(i % p) + (j % q) * p][ri, cj] - LU[ri, cj]).max() <= 1e-11 * np.abs(LU).max()
