"""GPU parity tests of the other side / op variants of trmm / hemm / symm on lower storage (SURVEY section 8(f) item 3;
slate_b200/csrc/solve.cu: trmm_lower_variant, hemm_symm_right_lower), through the host mirror of the reference API,
against (1) the golden vectors the UNMODIFIED reference wrote (tests/golden/make_golden.py blas3_variants) and (2) the
numpy restatement pinned to them (tests/test_oracle.py).

STATUS: written after round 2's GPU budget was spent.  The drivers are host-side compositions of launches that ARE validated
(the batched tile GEMM with an op on either operand, the diagonal-tile fill kernels, the one-block workspace of
trmm_left_lower) -- the shipped library's 179 kernels are bitwise the measured ones (scratch/sass_compare.py) -- and
their schedules are checked on the CPU (tests/test_blas3_variant_schedule.py), and the logic of THIS file (shapes, goldens,
tolerances) passes against an independent numpy stand-in for the host API (scratch/cpu_standin/check_gpu_test_logic.py: 604
cases) and the shipped host code passes all of them on the CPU over an emulated CUDA runtime (scratch/cpu_standin/fake_cudart.cc,
profiles/r02t_host_code_on_emulated_runtime.txt), but this file has NOT yet run on a B200.
It sorts last and is marked xfail(strict=False) for that reason alone: the tail of the first GPU run says whether the
cases XPASS (then the mark goes) without a first-run surprise hiding the 1 600 validated tests before it under `-x`."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o
from tests.gpu_util import NP

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="written after the round's GPU budget ended: first run on a B200 pending")]
EPS = float(np.finfo(np.float64).eps)
ALPHA = 3.141592653589793 + 1.414213562373095j
BETA = 2.718281828459045 + 1.732050807568877j


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _eps(t):
    return EPS if t in "dz" else float(np.finfo(np.float32).eps)


def _wide(t):
    return np.complex128 if t in "cz" else np.float64


@pytest.mark.parametrize("name,t,side,op,diag", [
    ("trmm_z_left_conj", "z", "L", "C", "N"), ("trmm_d_left_trans", "d", "L", "T", "U"),
    ("trmm_d_right", "d", "R", "N", "N"), ("trmm_z_right_trans", "z", "R", "T", "N"),
    ("trmm_z_right_conj", "z", "R", "C", "U")])
def test_trmm_variants_match_reference_golden(sl, golden_dir, name, t, side, op, diag):
    g = np.load(os.path.join(golden_dir, name + ".npz"))["out"]
    (m, n), nb = ((200, 70) if side == "L" else (70, 200)), 64
    A = sl.HermitianMatrix(m if side == "L" else n, nb, dtype=t).generate("rand", 42)      # its lower tiles are the triangle
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    sl.trmm(ALPHA if t == "z" else ALPHA.real, A, B, side=side, op=op, diag=diag)
    assert np.abs(B.to_host() - g).max() <= 3 * np.sqrt(max(m, n)) * EPS * 4 * np.abs(g).max()


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("side,op", [("L", "T"), ("L", "C"), ("R", "N"), ("R", "T"), ("R", "C")])
@pytest.mark.parametrize("diag", ["N", "U"])
@pytest.mark.parametrize("m,n,nb", [(200, 70, 64), (70, 200, 64), (512, 512, 128), (300, 1000, 256)])
def test_trmm_variants_match_oracle(sl, t, side, op, diag, m, n, nb):
    al = ALPHA if t in "cz" else ALPHA.real
    na = m if side == "L" else n
    A = sl.HermitianMatrix(na, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    sl.trmm(al, A, B, side=side, op=op, diag=diag)
    a = np.tril(o.generate("rand", na, na, 42, NP[t])).astype(_wide(t))
    b = o.generate("rand", m, n, 43, NP[t]).astype(_wide(t))
    ref = o.trmm(al, a, b, nb, unit=(diag == "U"), side=side, op=op)
    assert np.abs(B.to_host() - ref).max() <= 3 * np.sqrt(na) * _eps(t) * 4 * np.abs(ref).max()


def test_trmm_left_notrans_still_takes_the_validated_driver(sl):
    """side L, op N is trmm_left_lower as before (bitwise: the same launches)"""
    m, n, nb = 200, 70, 64
    out = []
    for _ in range(2):
        A = sl.HermitianMatrix(m, nb).generate("rand", 42)
        B = sl.Matrix(m, n, nb).generate("rand", 43)
        sl.trmm(ALPHA.real, A, B, side="L", op="N")
        out.append(B.to_host())
    ref = o.trmm(ALPHA.real, np.tril(o.generate("rand", m, m, 42)), o.generate("rand", m, n, 43), nb)
    assert np.array_equal(out[0], out[1])
    assert np.abs(out[0] - ref).max() <= 3 * np.sqrt(m) * EPS * 4 * np.abs(ref).max()


@pytest.mark.parametrize("name,routine,t,n", [("hemm_z_right", "hemm", "z", 192), ("hemm_d_right", "hemm", "d", 200),
                                              ("symm_z_right", "symm", "z", 192)])
def test_hemm_symm_right_match_reference_golden(sl, golden_dir, name, routine, t, n):
    g = np.load(os.path.join(golden_dir, name + ".npz"))["out"]
    nb, m = 64, 70
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(m, n, nb, dtype=t).generate("rand", 44)
    al, be = (ALPHA, BETA) if t == "z" else (ALPHA.real, BETA.real)
    (sl.hemm if routine == "hemm" else sl.symm)(al, A, B, be, C, side="R")
    assert np.abs(C.to_host() - g).max() <= 8 * np.sqrt(n) * EPS * np.abs(g).max()


@pytest.mark.parametrize("routine", ["hemm", "symm"])
@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("m,n,nb,beta", [(70, 192, 64, None), (10, 1000, 128, 1.0), (300, 512, 256, 0.0), (7, 300, 64, -0.5)])
def test_hemm_symm_right_match_oracle(sl, routine, t, m, n, nb, beta):
    al = ALPHA if t in "cz" else ALPHA.real
    be = (BETA if t in "cz" else BETA.real) if beta is None else beta
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(m, n, nb, dtype=t).generate("rand", 44)
    (sl.hemm if routine == "hemm" else sl.symm)(al, A, B, be, C, side="R")
    a = np.tril(o.generate("rand", n, n, 42, NP[t])).astype(_wide(t))
    b, c = (o.generate("rand", m, n, seed, NP[t]).astype(_wide(t)) for seed in (43, 44))
    ref = (o.hemm if routine == "hemm" else o.symm)(al, a, b, be, c, nb, side="R")
    full = o.he_full(a) if routine == "hemm" else o.sy_full(a)
    scale = np.abs(b) @ np.abs(full)
    assert (np.abs(C.to_host() - ref) <= 4 * np.sqrt(n) * _eps(t) * (abs(al) * scale + abs(be) * np.abs(c) + 1.0)).all()


def test_side_argument_left_forwards_and_bad_shapes_are_rejected(sl):
    n, nb, nrhs = 192, 64, 70
    res = []
    for side in (None, "L"):
        A = sl.HermitianMatrix(n, nb, dtype="z").generate("rand", 42)
        B = sl.Matrix(n, nrhs, nb, dtype="z").generate("rand", 43)
        C = sl.Matrix(n, nrhs, nb, dtype="z").generate("rand", 44)
        if side is None:
            sl.hemm(ALPHA, A, B, BETA, C)
        else:
            sl.hemm(ALPHA, A, B, BETA, C, side=side)
        res.append(C.to_host())
    assert np.array_equal(res[0], res[1])
    A = sl.HermitianMatrix(64, 32); B = sl.Matrix(64, 8, 32); C = sl.Matrix(64, 8, 32)
    with pytest.raises(sl.SB200Error):
        sl.hemm(1.0, A, B, 0.0, C, side="R")               # Side::Right needs B.n == A.n
    with pytest.raises(sl.SB200Error):
        sl.trmm(1.0, A, B, side="R")
    with pytest.raises(sl.SB200Error):
        sl.trmm(1.0, A, B, uplo="U")                       # lower storage only


# ---------------------------------------------------------------------------------------------------------------
# slate::trsm at matrix level (sb200_trsm_mat_*): Side::Left through the sweep potrs / getrs run (tri_sweep), Side::Right
# through tri_sweep_right; posv / gesv (slate::posv, slate::gesv: factor, then solve when info == 0)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,t,side,op,diag", [
    ("trsm_d", "d", "L", "N", "N"),                       # the round-1 golden: slate::triangular_solve, m = 256, n = 128
    ("trsm_z_left_conj", "z", "L", "C", "N"), ("trsm_d_left_trans", "d", "L", "T", "U"),
    ("trsm_d_right", "d", "R", "N", "N"), ("trsm_z_right_trans", "z", "R", "T", "U"),
    ("trsm_z_right_conj", "z", "R", "C", "N")])
def test_trsm_matrix_level_matches_reference_golden(sl, golden_dir, name, t, side, op, diag):
    g = np.load(os.path.join(golden_dir, name + ".npz"))["out"]
    (m, n), nb = ((256, 128) if name == "trsm_d" else (200, 70) if side == "L" else (70, 200)), 64
    A = sl.HermitianMatrix(m if side == "L" else n, nb, dtype=t).generate("rand_dominant", 42)
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    sl.trsm(ALPHA if t == "z" else ALPHA.real, A, B, side=side, op=op, diag=diag)
    # unit-diagonal solves with this triangle grow (|X| up to 7e9): the bound is relative to the largest entry
    assert np.abs(B.to_host() - g).max() <= 1e-10 * np.abs(g).max()


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("side", ["L", "R"])
@pytest.mark.parametrize("op", ["N", "T", "C"])
@pytest.mark.parametrize("m,n,nb", [(200, 70, 64), (70, 200, 64), (512, 512, 128), (300, 1000, 256)])
def test_trsm_matrix_level_lower_vs_oracle(sl, t, side, op, m, n, nb):
    al = ALPHA if t in "cz" else ALPHA.real
    na = m if side == "L" else n
    A = sl.HermitianMatrix(na, nb, dtype=t).generate("rand_dominant", 42)
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    sl.trsm(al, A, B, side=side, op=op)
    a = np.tril(o.generate("rand_dominant", na, na, 42, NP[t])).astype(_wide(t))
    b = o.generate("rand", m, n, 43, NP[t]).astype(_wide(t))
    ref = o.trsm(al, a, b, nb, side=side, lower=True, op=op)
    assert np.abs(B.to_host() - ref).max() <= 200 * _eps(t) * np.abs(ref).max()        # the bound of the tile-level trsm tests


@pytest.mark.parametrize("side,uplo,op,diag", [("L", "U", "N", "N"), ("L", "U", "C", "N"), ("R", "U", "N", "N"), ("R", "U", "T", "N"),
                                               ("L", "L", "N", "U"), ("R", "L", "N", "U"), ("R", "L", "C", "U")])
@pytest.mark.parametrize("t", ["d", "z"])
def test_trsm_matrix_level_on_an_lu_factor(sl, t, side, uplo, op, diag):
    """A general square matrix as the triangle holder (what getrs does with L and U): upper triangle, and the unit lower
    one.  Checked through the backward error || op(T) X - alpha B ||_1 / (|| T ||_1 || X ||_1) (Left; Right likewise) --
    the reference tester's residual (test/test_trsm.cc:160-184: <= 3 eps) normalised by || X || instead of N so that it
    stays meaningful for the unit-diagonal solves, which grow to 1e7 ... 1e14 on this triangle; the oracle's own sweep
    reaches 1.9 eps, the bound here is 16 eps -- and against the oracle where the solve does not grow."""
    n, nb, nrhs = 300, 64, 70
    A = sl.Matrix(n, n, nb, dtype=t).generate("rand_dominant", 42)
    shape = (n, nrhs) if side == "L" else (nrhs, n)
    B = sl.Matrix(*shape, nb, dtype=t).generate("rand", 43)
    al = ALPHA if t == "z" else ALPHA.real
    sl.trsm(al, A, B, side=side, uplo=uplo, op=op, diag=diag)
    X = B.to_host()
    a = o.generate("rand_dominant", n, n, 42, NP[t])
    tri = np.tril(a) if uplo == "L" else np.triu(a)
    if diag == "U":
        np.fill_diagonal(tri, 1.0)
    M = {"N": tri, "T": tri.T, "C": tri.conj().T}[op]
    b = o.generate("rand", *shape, 43, NP[t])
    R = (M @ X - al * b) if side == "L" else (X @ M - al * b)
    resid = np.abs(R).sum(axis=0).max() / (np.abs(M).sum(axis=0).max() * np.abs(X).sum(axis=0).max())
    assert resid <= 16 * EPS
    if diag == "N":
        ref = o.trsm(al, a, b, nb, side=side, lower=(uplo == "L"), op=op)
        assert np.abs(X - ref).max() <= 200 * EPS * np.abs(ref).max()


def test_posv_and_gesv_are_factor_then_solve(sl, golden_dir):
    """slate::posv / slate::gesv (src/posv.cc:80-94, src/gesv.cc:95-109) against the reference's golden solutions; a
    matrix that is not positive definite returns its info and leaves B alone."""
    g = np.load(os.path.join(golden_dir, "posv_d.npz"))
    n, nb = 300, 128
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    assert sl.posv(A, B) == 0 == int(g["info"])
    assert np.abs(B.to_host() - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()
    g = np.load(os.path.join(golden_dir, "gesv_d.npz"))
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    piv, info = sl.lu_solve(A, B)
    assert info == 0 == int(g["info"]) and len(piv) == 3
    assert np.abs(B.to_host() - g["out"]).max() <= 1e-10 * np.abs(g["out"]).max()
    H = sl.HermitianMatrix(n, nb).generate("rand", 42)                 # rand is not positive definite
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    b0 = B.to_host()
    info = sl.chol_solve(H, B)
    _, iref = o.potrf(o.he_full(np.tril(o.generate("rand", n, n, 42))), nb)
    assert info == iref > 0
    assert np.array_equal(B.to_host(), b0)


# ---------------------------------------------------------------------------------------------------------------
# slate::gemm with (conjugate-)transposed views (sb200_gemm_op_*: solve.cu gemm_ops)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,t,opa,opb", [("gemm_d_tn", "d", "T", "N"), ("gemm_d_nt", "d", "N", "T"), ("gemm_z_cn", "z", "C", "N"),
                                           ("gemm_z_tc", "z", "T", "C"), ("gemm_z_nc", "z", "N", "C")])
def test_gemm_transposed_views_match_reference_golden(sl, golden_dir, name, t, opa, opb):
    g = np.load(os.path.join(golden_dir, name + ".npz"))["out"]
    m, n, k, nb = 150, 200, 100, 64
    A = sl.Matrix(*((m, k) if opa == "N" else (k, m)), nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(*((k, n) if opb == "N" else (n, k)), nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(m, n, nb, dtype=t).generate("rand", 44)
    al, be = (ALPHA, BETA) if t == "z" else (ALPHA.real, BETA.real)
    sl.gemm(al, A, B, be, C, opA=opa, opB=opb)
    assert np.abs(C.to_host() - g).max() <= 3 * np.sqrt(k) * EPS * 4 * np.abs(g).max()


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("opa", ["N", "T", "C"])
@pytest.mark.parametrize("opb", ["N", "T", "C"])
@pytest.mark.parametrize("m,n,k,nb", [(150, 200, 100, 64), (512, 512, 512, 128), (70, 10, 300, 64), (300, 1000, 260, 256)])
def test_gemm_transposed_views_vs_oracle_and_tester_check(sl, t, opa, opb, m, n, k, nb):
    al, be = (ALPHA, BETA) if t in "cz" else (ALPHA.real, BETA.real)
    sa, sb = ((m, k) if opa == "N" else (k, m)), ((k, n) if opb == "N" else (n, k))
    A = sl.Matrix(*sa, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(*sb, nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(m, n, nb, dtype=t).generate("rand", 44)
    sl.gemm(al, A, B, be, C, opA=opa, opB=opb)
    a, b, c0 = (o.generate("rand", *shape, seed, NP[t]).astype(_wide(t)) for shape, seed in ((sa, 42), (sb, 43), ((m, n), 44)))
    ref = o.gemm(al, a, b, be, c0, nb, opa=opa, opb=opb)
    out = C.to_host()
    assert np.abs(out - ref).max() <= 3 * np.sqrt(k) * _eps(t) * 4 * np.abs(ref).max()
    view = {"N": lambda x: x, "T": lambda x: x.T, "C": lambda x: x.conj().T}
    assert o.gemm_check(al, view[opa](a), view[opb](b), be, c0, out.astype(_wide(t))) <= 3 * _eps(t)     # test/test_gemm.cc:192-208


# ---------------------------------------------------------------------------------------------------------------
# More of the reference's OWN tester through the shim (oracle/_ref/tester_sb200, see tests/test_reference_tester_gpu.py):
# routines of SURVEY section 8(f) and BASELINE configs[4] whose SLATE drivers reach these kernels through the
# blas::batch::gemm / trsm / herk, lapack::potrf and slate::device::* seams under Target::Devices.  The tester's own
# residual checks decide; routines it can only check against ScaLAPACK report "no check" and prove that the Devices
# driver runs to completion on these kernels.  (First run pending like the rest of this file; each run is a subprocess.)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("routine,extra,nrows", [
    ("posv_mixed",  ["--type", "d", "--dim", "2048", "--nb", "256"], 1),
    ("gesv_mixed",  ["--type", "d", "--dim", "2048", "--nb", "256"], 1),
    ("potrs",       ["--type", "d,z", "--dim", "2048", "--nb", "256"], 2),
    ("getrs",       ["--type", "d", "--dim", "2048", "--nb", "256"], 1),
    ("getrf_nopiv", ["--type", "d", "--dim", "2048", "--nb", "256", "--matrix", "rand_dominant"], 1),
    ("gesv_nopiv",  ["--type", "d", "--dim", "2048", "--nb", "256", "--matrix", "rand_dominant"], 1),
    ("her2k",       ["--type", "d,z", "--dim", "1024", "--nb", "256"], 2),
    ("syr2k",       ["--type", "d,z", "--dim", "768", "--nb", "256"], 2),
    ("gemmA",       ["--type", "d", "--dim", "1024x64x1024", "--nb", "256"], 1),
    ("tzadd",       ["--type", "d,z", "--dim", "1000x800", "--nb", "256", "--uplo", "l,u"], 4),
    ("tzscale",     ["--type", "d,z", "--dim", "1000x800", "--nb", "256", "--uplo", "l,u"], 4),
    ("tzcopy",      ["--type", "d,z", "--dim", "1000x800", "--nb", "256", "--uplo", "l,u"], 4),
    ("tzset",       ["--type", "d,z", "--dim", "1000x800", "--nb", "256", "--uplo", "l,u"], 4),
    ("scale_row_col", ["--type", "d,z", "--dim", "1000x800", "--nb", "256"], 2),
])
def test_more_reference_tester_routines_on_our_kernels(routine, extra, nrows):
    from tests.test_reference_tester_gpu import run_tester
    rc, text, rows = run_tester("--target", "d", "--origin", "d", "--check", "y", "--ref", "n", *extra, routine)
    assert rc == 0, text[-3000:]
    assert len(rows) >= nrows, text[-3000:]
    for r in rows:
        assert ("pass" in r or "no check" in r) and "FAILED" not in r and "failed" not in r, r
    assert "All tests passed" in text, text[-2000:]


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("n,nb,nrhs", [(300, 64, 10), (1024, 256, 130)])
def test_gesv_nopiv_is_factor_then_two_sweeps(sl, t, n, nb, nrhs):
    """slate::gesv_nopiv (src/gesv_nopiv.cc; getrs_nopiv = the unit-lower and the upper sweep of getrs without the
    permutation, src/getrs_nopiv.cc:20-53) on a diagonally dominant matrix: the tester's solve residual."""
    A = sl.Matrix(n, n, nb, dtype=t).generate("rand_dominant", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    assert sl.lu_solve_nopiv(A, B) == 0
    a = o.generate("rand_dominant", n, n, 42, NP[t]).astype(np.float64)
    b = o.generate("rand", n, nrhs, 43, NP[t]).astype(np.float64)
    X = B.to_host().astype(np.float64)
    assert o.solve_residual(a, X, b) <= 25 * _eps(t)                     # test/test_gesv.cc:371-377
    LUo, info = o.getrf_nopiv(a, nb)
    assert info == 0
    Xo = o.tri_sweep(np.triu(LUo), o.tri_sweep(np.tril(LUo), b, nb, lower=True, unit=True), nb, lower=False)
    assert np.abs(X - Xo).max() <= 200 * _eps(t) * np.abs(Xo).max()


# ---------------------------------------------------------------------------------------------------------------
# herk / her2k / syrk / syr2k handed (conjugate-)transposed views (sb200_{herk,her2k,syrk,syr2k}_op_*: rank_update_trans)
# ---------------------------------------------------------------------------------------------------------------
def _rank_update(sl, routine, t, n, k, nb):
    cplx = t in "cz"
    al, be = (ALPHA, BETA) if cplx else (ALPHA.real, BETA.real)
    A = sl.Matrix(k, n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(k, n, nb, dtype=t).generate("rand", 43)
    C = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 44)
    a0, b0 = (o.generate("rand", k, n, seed, NP[t]).astype(_wide(t)) for seed in (42, 43))
    c = np.tril(o.generate("rand", n, n, 44, NP[t]).astype(_wide(t)))
    if routine == "herk":
        sl.herk(al.real, A, be.real, C, op="C")
        ref = o.herk(al.real, a0.conj().T, be.real, c, nb)
    elif routine == "her2k":
        sl.her2k(al, A, B, be.real, C, op="C")
        ref = o.her2k(al, a0.conj().T, b0.conj().T, be.real, c, nb)
    elif routine == "syrk":
        sl.syrk(al, A, be, C, op="T")
        ref = o.syrk(al, a0.T, be, c, nb)
    else:
        sl.syr2k(al, A, B, be, C, op="T")
        ref = o.syr2k(al, a0.T, b0.T, be, c, nb)
    return np.tril(C.to_host()), np.tril(ref)


@pytest.mark.parametrize("name,routine,t", [("herk_z_conj", "herk", "z"), ("herk_d_trans", "herk", "d"), ("her2k_z_conj", "her2k", "z"),
                                            ("syrk_z_trans", "syrk", "z"), ("syr2k_z_trans", "syr2k", "z")])
def test_rank_updates_with_transposed_views_match_reference_golden(sl, golden_dir, name, routine, t):
    g = np.tril(np.load(os.path.join(golden_dir, name + ".npz"))["out"])
    out, _ = _rank_update(sl, routine, t, 200, 100, 64)
    assert np.abs(out - g).max() <= 64 * EPS * np.abs(g).max()


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("routine", ["herk", "her2k", "syrk", "syr2k"])
@pytest.mark.parametrize("n,k,nb", [(200, 100, 64), (512, 512, 128), (300, 1000, 256), (70, 10, 64)])
def test_rank_updates_with_transposed_views_vs_oracle(sl, t, routine, n, k, nb):
    out, ref = _rank_update(sl, routine, t, n, k, nb)
    assert np.abs(out - ref).max() <= 3 * np.sqrt(k) * _eps(t) * 4 * np.abs(ref).max()
    if routine in ("herk", "her2k") and t in "cz":
        assert np.all(np.diag(out).imag == 0)                      # the Hermitian diagonal is stored real


def test_rank_update_op_arguments_are_checked(sl):
    A = sl.Matrix(32, 64, 32, dtype="z"); C = sl.HermitianMatrix(64, 32, dtype="z")
    with pytest.raises(sl.SB200Error):
        sl.herk(1.0, A, 0.0, C, op="T")                            # complex herk takes a conjugate-transposed view only
    with pytest.raises(sl.SB200Error):
        sl.syrk(1.0, A, 0.0, C, op="C")                            # complex syrk a transposed one
    with pytest.raises(sl.SB200Error):
        sl.herk(1.0, sl.Matrix(32, 48, 32, dtype="z"), 0.0, C, op="C")       # stored k x n: n has to match C


# ---------------------------------------------------------------------------------------------------------------
# getrs handed a (conjugate-)transposed view: op(A) X = B with the factors of A (sb200_getrs_op: getrs_trans_t)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,t,n,op", [("gesv_d_trans", "d", 300, "T"), ("gesv_z_conj", "z", 200, "C")])
def test_getrs_transposed_matches_reference_golden(sl, golden_dir, name, t, n, op):
    g = np.load(os.path.join(golden_dir, name + ".npz"))["out"]
    nb, nrhs = 64, 70
    A = sl.Matrix(n, n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    sl.getrs(A, piv, B, op=op)
    X = B.to_host()
    assert np.abs(X - g).max() <= 1e-9 * np.abs(g).max()
    a = o.generate("rand", n, n, 42, NP[t])
    M = a.T if op == "T" else a.conj().T
    assert o.solve_residual(M, X, o.generate("rand", n, nrhs, 43, NP[t])) <= 25 * EPS        # test/test_gesv.cc:371-377


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("op", ["T", "C"])
@pytest.mark.parametrize("n,nb,nrhs", [(1024, 256, 10), (700, 128, 130), (300, 64, 3)])
def test_getrs_transposed_vs_oracle_and_tester_residual(sl, t, op, n, nb, nrhs):
    A = sl.Matrix(n, n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    piv, info = sl.getrf(A)
    assert info == 0
    LU = A.to_host().astype(_wide(t))
    sl.getrs(A, piv, B, op=op)
    X = B.to_host().astype(_wide(t))
    a = o.generate("rand", n, n, 42, NP[t]).astype(_wide(t))
    b = o.generate("rand", n, nrhs, 43, NP[t]).astype(_wide(t))
    M = a.T if (op == "T" or t in "ds") else a.conj().T
    assert o.solve_residual(M, X, b) <= 25 * _eps(t)
    if t in "dz":
        Xo = o.getrs(LU, piv, b, nb, op=op)                 # the same factors through the oracle's sweeps
        assert np.abs(X - Xo).max() <= 1e-9 * np.abs(Xo).max()


# ---------------------------------------------------------------------------------------------------------------
# slate::norm at matrix level, all four norms (sb200_norm_*: per-tile kernels + host combination, solve.cu norm_mat)
# ---------------------------------------------------------------------------------------------------------------
def test_norms_match_reference_golden(sl, golden_dir):
    g = np.load(os.path.join(golden_dir, "norms_d.npz"))["out"]           # max, one, inf, fro of rand 200 x 136, nb = 64
    A = sl.Matrix(200, 136, 64).generate("rand", 42)
    assert sl.norm("max", A) == g[0]
    assert abs(sl.norm("one", A) - g[1]) <= 8 * EPS * g[1]
    assert abs(sl.norm("inf", A) - g[2]) <= 8 * EPS * g[2]
    assert abs(sl.norm("fro", A) - g[3]) <= 64 * EPS * g[3]


def _np_norms(F):
    a = np.abs(F.astype(np.complex128 if np.iscomplexobj(F) else np.float64))
    return {"max": a.max(), "one": a.sum(axis=0).max(), "inf": a.sum(axis=1).max(), "fro": np.sqrt((a * a).sum())}


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("m,n,nb", [(200, 136, 64), (1000, 800, 256), (64, 64, 64), (50, 300, 128)])
def test_general_norms_vs_numpy(sl, t, m, n, nb):
    A = sl.Matrix(m, n, nb, dtype=t).generate("rand", 42)
    ref = _np_norms(o.generate("rand", m, n, 42, NP[t]))
    for kind, r in ref.items():
        v = sl.norm(kind, A)
        assert abs(v - r) <= (1 if kind == "max" else 4 * np.sqrt(max(m, n))) * _eps(t) * r, kind
    assert abs(sl.norm("inf", A) - sl.norm_inf(A)) <= 8 * max(m, n) * _eps(t) * ref["inf"]      # the validated one-kernel row sums


@pytest.mark.parametrize("t", ["d", "z", "c"])
@pytest.mark.parametrize("symmetric", [False, True])
@pytest.mark.parametrize("n,nb", [(200, 64), (1000, 256), (300, 512)])
def test_hermitian_and_symmetric_norms_vs_numpy(sl, t, symmetric, n, nb):
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 42)
    a = np.tril(o.generate("rand", n, n, 42, NP[t])).astype(_wide(t))
    full = o.sy_full(a) if symmetric else o.he_full(a)            # he_full takes the stored diagonal real, as the kernels do
    ref = _np_norms(full)
    for kind, r in ref.items():
        v = sl.norm(kind, A, symmetric=symmetric)
        assert abs(v - r) <= (1 if kind == "max" else 4 * np.sqrt(n)) * _eps(t) * r, kind
    assert abs(sl.norm("one", A, symmetric=symmetric) - sl.norm("inf", A, symmetric=symmetric)) <= 8 * n * _eps(t) * ref["one"]


def test_norm_propagates_nan_and_rejects_unknown_norms(sl):
    import torch
    h = np.asfortranarray(o.generate("rand", 100, 70, 1))
    h[37, 11] = np.nan
    A = sl.Matrix(100, 70, 32).from_host(h)
    for kind in ("max", "one", "inf", "fro"):
        assert np.isnan(sl.norm(kind, A)), kind
    with pytest.raises(sl.SB200Error):
        sl.norm("two", A)
