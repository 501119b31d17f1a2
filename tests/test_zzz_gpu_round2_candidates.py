"""GPU parity tests of the opt-in round-2 candidates (each behind an environment switch, default behaviour untouched):

* SB200_TILE_FUSED=1|2 -- one-launch tile Cholesky (slate_b200/csrc/potrf_tile_fused.cu), through the C ABI
  (`sb200_potrf_tile_d`, the lapack::potrf seam: src/internal/internal_potrf.cc:57-81) and through the driver.
* SB200_TRSM_FUSED bit 0 -- one-launch Cholesky panel solve B <- alpha B L^-T, bit 1 -- one-launch LU row solve
  B <- alpha L^-1 B (same file), through `sb200_trsm_batched_*` (the blas::batch::trsm seam:
  src/internal/internal_trsm.cc:132-262) and through the drivers.

* `potrf(A, in_local=..., out_local=...)` -- input AND output stream between pinned host memory and the device in chunks
  of block columns while the factorisation runs (`sb200_potrf_stream_*`, csrc/runtime.cu): bitwise the default factor.

* the LU base-block kernels of csrc/getrf_base_v3.cu (per-column exchange as {data | generation tag} 8-byte words, rows
  in registers, 32-column updates folded into the next launch) against their fallbacks: identical pivots.

* the 64 x 64 diagonal-block Cholesky with one rsqrt per column (default) and in the divided form (SB200_DIAG_RSQRT=0).

* SB200_GEMM_BT=1 -- dgemm (SUMMA) with the B row panel transposed once per step so that the multiply runs as 'N','T'
  (both operands through TMA bulk copies): bitwise the default C.

* `her2k`: C = alpha A B^H + conj(alpha) B A^H + beta C on the herk skeleton (two batched launches
  per step), against the reference's golden output and the oracle.

* `Matrix.from_scalapack / to_scalapack`: ScaLAPACK-style local array <-> HBM tile pool.

* `getrf_nopiv`: the getrf drivers with the pivot search compiled out of the base kernel.

Validated on a B200 in round 2 (gpurun call r2b: 357 passed; profiles/r02b_pytest_gpu_tail.txt); the winners are the
defaults in csrc/common.cuh (SW_* table)."""
import os

import numpy as np
import pytest

from oracle import slate_oracle as o
from tests.gpu_util import GETRF_TOL

pytestmark = [pytest.mark.gpu]
EPS = float(np.finfo(np.float64).eps)


@pytest.fixture(scope="module")
def sl():
    import torch
    torch.cuda.set_device(0)
    import slate_b200.host as sl_
    return sl_


def _potrf_tile(A, n, lda=None, t="d"):
    from tests.gpu_util import DevTiles, dev_zeros, fn, stream, sync, c_int, c_i64, c_ptr, NP
    lda = lda or n
    buf = np.full((lda, n), 7.25, dtype=NP[t])
    buf[:n, :] = A
    dA = DevTiles([buf])
    info = dev_zeros(1, np.int32)
    f = fn(f"sb200_potrf_tile_{t}", [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr])
    assert f(ord("L"), n, dA.t[0].data_ptr(), lda, info.data_ptr(), None, stream()) == 0
    sync()
    return dA.get()[0], int(info.cpu()[0])


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("n,lda", [(128, 128), (512, 512), (448, 512), (500, 500), (130, 136), (1000, 1024), (1024, 1024), (65, 65)])
def test_fused_tile_cholesky_vs_lapack(monkeypatch, variant, n, lda, t):
    monkeypatch.setenv("SB200_TILE_FUSED", variant)
    rng = np.random.default_rng(5)
    G = rng.random((n, n))
    A = (G @ G.T + n * np.eye(n)).astype(np.float64 if t == "d" else np.float32)
    out, info = _potrf_tile(A, n, lda, t)
    assert info == 0
    ref = np.linalg.cholesky(A.astype(np.float64))
    eps = EPS if t == "d" else float(np.finfo(np.float32).eps)
    assert np.abs(np.tril(out[:n, :]) - ref).max() <= 50 * eps * np.abs(ref).max()
    assert np.array_equal(np.triu(out[:n, :], 1), np.triu(A, 1))          # strict upper triangle untouched
    assert np.all(out[n:, :] == 7.25)                                    # rows beyond n (lda > n) untouched


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("n,bad", [(128, 70), (512, 300), (512, 0), (512, 511), (500, 64), (256, 63)])
def test_fused_tile_cholesky_reports_first_bad_minor(monkeypatch, variant, n, bad):
    monkeypatch.setenv("SB200_TILE_FUSED", variant)
    A = np.eye(n) * 4.0
    A[bad, bad] = -1.0
    if bad + 10 < n:
        A[bad + 10, bad + 10] = -2.0                                     # a later failure must not win
    _, info = _potrf_tile(A, n)
    assert info == bad + 1


@pytest.mark.parametrize("variant", ["1", "2"])
def test_fused_tile_cholesky_same_result_as_default_path(monkeypatch, variant):
    """Same blocked algorithm, different summation order inside a tile: a few ulp apart, not bitwise."""
    n = 512
    rng = np.random.default_rng(9)
    G = rng.random((n, n))
    A = G @ G.T + n * np.eye(n)
    monkeypatch.delenv("SB200_TILE_FUSED", raising=False)
    base, info0 = _potrf_tile(A, n)
    monkeypatch.setenv("SB200_TILE_FUSED", variant)
    out, info1 = _potrf_tile(A, n)
    assert info0 == info1 == 0
    assert np.abs(np.tril(out) - np.tril(base)).max() <= 32 * EPS * np.abs(base).max()


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("n,nb", [(2048, 512), (1000, 256), (1100, 512), (4096, 1024)])
def test_potrf_driver_with_fused_tile(sl, monkeypatch, variant, n, nb):
    monkeypatch.setenv("SB200_TILE_FUSED", variant)
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 11)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    G = o.generate("rand_dominant", n, n, 11)
    Af = np.tril(G) + np.tril(G, -1).T
    Lo, info = o.potrf(Af, nb)
    assert info == 0
    assert np.abs(L - Lo).max() <= 64 * EPS * np.abs(Lo).max()
    assert np.abs(L @ L.T - Af).max() <= 64 * EPS * np.abs(Af).max()


def test_potrf_driver_with_fused_tile_info(sl, monkeypatch):
    monkeypatch.setenv("SB200_TILE_FUSED", "1")
    n, nb = 1024, 256
    S = np.eye(n) * 3.0
    S[700, 700] = -1.0
    A = sl.HermitianMatrix(n, nb); A.from_host(np.asfortranarray(S))
    assert sl.potrf(A) == 701


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("n,nb", [(2048, 512), (1000, 128)])
def test_posv_mixed_with_fused_tile(sl, monkeypatch, variant, n, nb):
    """The FP32 factorisation of posv_mixed takes the FP32 instance of the fused tile kernel."""
    monkeypatch.setenv("SB200_TILE_FUSED", variant)
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    info, it, _ = sl.posv_mixed(A, B, X)
    assert info == 0 and 0 <= it <= 30
    a = o.generate("rand_dominant", n, n, 42); b = o.generate("rand", n, 10, 43)
    assert o.solve_residual(o.he_full(a), X.to_host(), b) <= 25 * EPS            # test/test_posv.cc:336-342


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("m,n", [(512, 512), (300, 512), (512, 448), (100, 130), (64, 500), (1, 65), (512, 1024)])
def test_fused_panel_trsm_right_lower_trans(monkeypatch, t, m, n):
    """Right / Lower / Trans / NonUnit (the potrf panel solve) through the C ABI, vs the oracle's tile solve; the layout
    'R' call of the same routine is the Left / Upper / NoTrans problem on the transposed storage and must agree too."""
    from tests.gpu_util import DevTiles, fn, scal, stream, rng_tiles, NP, SC, c_int, c_i64, c_ptr
    monkeypatch.setenv("SB200_TRSM_FUSED", "1")
    rng = np.random.default_rng(4)
    batch = 3
    T = (rng.random((n, n)) / n + np.eye(n) * (1 + rng.random(n))).astype(NP[t])
    B = rng_tiles(rng, batch, m, n, t)
    alpha = 0.7
    ref = [o.trsm_tile("R", "L", "T", "N", alpha, T.astype(np.float64), b.astype(np.float64)) for b in B]
    dT, dB = DevTiles([T]), DevTiles(B)
    f = fn(f"sb200_trsm_batched_{t}", [c_int] * 5 + [c_i64, c_i64, SC[t], c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr])
    assert f(ord("C"), ord("R"), ord("L"), ord("T"), ord("N"), m, n, scal(t, alpha), dT.t[0].data_ptr(), n,
             dB.p, m, batch, None, stream()) == 0
    eps = EPS if t == "d" else float(np.finfo(np.float32).eps)
    for x, r in zip(dB.get(), ref):
        assert np.abs(x - r).max() <= 200 * eps * np.abs(r).max()


@pytest.mark.parametrize("n,nb", [(2048, 512), (1100, 512), (1000, 256)])
def test_potrf_driver_with_fused_tile_and_panel_solve(sl, monkeypatch, n, nb):
    monkeypatch.setenv("SB200_TILE_FUSED", "1")
    monkeypatch.setenv("SB200_TRSM_FUSED", "1")
    A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 11)
    assert sl.potrf(A) == 0
    L = np.tril(A.to_host())
    G = o.generate("rand_dominant", n, n, 11)
    Af = np.tril(G) + np.tril(G, -1).T
    Lo, info = o.potrf(Af, nb)
    assert info == 0
    assert np.abs(L - Lo).max() <= 64 * EPS * np.abs(Lo).max()
    assert np.abs(L @ L.T - Af).max() <= 64 * EPS * np.abs(Af).max()


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("diag", ["U", "N"])
@pytest.mark.parametrize("m,n", [(512, 512), (512, 300), (448, 512), (130, 100), (500, 64), (65, 1), (1024, 512), (256, 700)])
def test_fused_row_trsm_left_lower_notrans(monkeypatch, t, diag, m, n):
    """Left / Lower / NoTrans (the LU row solve U(k,j) = L_kk^-1 A(k,j), unit diagonal in getrf) through the C ABI."""
    from tests.gpu_util import DevTiles, fn, scal, stream, rng_tiles, NP, SC, c_int, c_i64, c_ptr
    monkeypatch.setenv("SB200_TRSM_FUSED", "2")
    rng = np.random.default_rng(4)
    batch = 3
    T = (rng.random((m, m)) / m + np.eye(m) * (1 + rng.random(m))).astype(NP[t])
    B = rng_tiles(rng, batch, m, n, t)
    alpha = 0.7
    ref = [o.trsm_tile("L", "L", "N", diag, alpha, T.astype(np.float64), b.astype(np.float64)) for b in B]
    dT, dB = DevTiles([T]), DevTiles(B)
    f = fn(f"sb200_trsm_batched_{t}", [c_int] * 5 + [c_i64, c_i64, SC[t], c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr])
    assert f(ord("C"), ord("L"), ord("L"), ord("N"), ord(diag), m, n, scal(t, alpha), dT.t[0].data_ptr(), m,
             dB.p, m, batch, None, stream()) == 0
    eps = EPS if t == "d" else float(np.finfo(np.float32).eps)
    for x, r in zip(dB.get(), ref):
        assert np.abs(x - r).max() <= 200 * eps * np.abs(r).max()


@pytest.mark.parametrize("dist", ["0", "1"])
@pytest.mark.parametrize("m,n,nb", [(1024, 1024, 256), (2048, 2048, 512), (700, 300, 128), (300, 700, 128), (1100, 1100, 512)])
def test_getrf_with_fused_row_solve_identical_pivots(sl, monkeypatch, m, n, nb, dist):
    monkeypatch.setenv("SB200_TRSM_FUSED", "2")
    monkeypatch.setenv("SB200_GETRF_DIST", dist)
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, info_o = o.getrf(A0, nb, 32)
    assert info == info_o == 0
    assert piv == pivo, "pivot vectors differ from the oracle's"
    assert np.abs(A.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()


def test_gesv_mixed_with_all_fused_candidates(sl, monkeypatch):
    monkeypatch.setenv("SB200_TRSM_FUSED", "3")
    monkeypatch.setenv("SB200_TILE_FUSED", "1")
    n, nb = 2048, 512
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    info, it, piv, tm = sl.gesv_mixed(A, B, X)
    assert info == 0 and 0 <= it <= 30
    assert o.solve_residual(o.generate("rand", n, n, 42), X.to_host(), o.generate("rand", n, 10, 43)) <= 25 * EPS


@pytest.mark.parametrize("chunk", ["1", "2", "8"])
@pytest.mark.parametrize("t,n,nb", [("d", 2048, 256), ("d", 1100, 128), ("d", 4096, 512), ("z", 1024, 128), ("s", 1024, 256), ("d", 300, 512)])
def test_potrf_streaming_input_is_bitwise_the_default_factor(sl, monkeypatch, chunk, t, n, nb):
    import torch
    monkeypatch.setenv("SB200_STREAM_CHUNK", chunk)
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand_dominant", 7)
    tdt = {"d": torch.float64, "s": torch.float32, "z": torch.complex128}[t]
    nelem = A.local_tiles * nb * nb
    hin = torch.empty(nelem, dtype=tdt).pin_memory()
    hout = torch.zeros(nelem, dtype=tdt).pin_memory()
    ref = torch.empty(nelem, dtype=tdt).pin_memory()
    A.to_host_local(hin)
    assert sl.potrf(A) == 0
    A.to_host_local(ref)
    B = sl.HermitianMatrix(n, nb, dtype=t)                     # content irrelevant: overwritten by the stream
    assert sl.potrf(B, in_local=hin, out_local=hout) == 0
    got = torch.empty(nelem, dtype=tdt).pin_memory()
    B.to_host_local(got)
    assert torch.equal(got.view(torch.uint8), ref.view(torch.uint8)), "device factor differs from the default path"
    # the streamed-out copy holds every block column as it was when it became final == the final factor
    assert torch.equal(hout.view(torch.uint8), ref.view(torch.uint8)), "streamed-out factor differs"


def test_potrf_streaming_reports_info(sl):
    import torch
    n, nb = 1024, 128
    S = np.eye(n) * 3.0
    S[700, 700] = -1.0
    A = sl.HermitianMatrix(n, nb); A.from_host(np.asfortranarray(S))
    hin = torch.empty(A.local_tiles * nb * nb, dtype=torch.float64).pin_memory()
    A.to_host_local(hin)
    B = sl.HermitianMatrix(n, nb)
    assert sl.potrf(B, in_local=hin) == 701


# LU base-block kernels (csrc/getrf_base_v3.cu): the default is the register-resident kernel with the 32-column updates
# folded into the next launch; the others are its fallbacks (tall panels, getrf_nopiv, SB200_PANEL=1) and must give the
# same pivots and the same factor up to rounding.
BASE_VARIANTS = {"default": {}, "no_fused_update": {"SB200_PANEL_FUSE": "0"}, "shared_memory_rows": {"SB200_PANEL_V4": "0"},
                 "cooperative_barrier": {"SB200_PANEL_V3": "0"}}


def _set_variant(monkeypatch, variant):
    for k in ("SB200_PANEL_FUSE", "SB200_PANEL_V4", "SB200_PANEL_V3"):
        monkeypatch.delenv(k, raising=False)
    for k, v in BASE_VARIANTS[variant].items():
        monkeypatch.setenv(k, v)


@pytest.mark.parametrize("variant", list(BASE_VARIANTS))
@pytest.mark.parametrize("dist", ["0", "1"])
@pytest.mark.parametrize("panel", ["1", "2"])
@pytest.mark.parametrize("m,n,nb", [(1024, 1024, 256), (700, 700, 128), (2048, 2048, 512), (700, 300, 128), (300, 700, 128),
                                    (1100, 1100, 512), (1000, 1000, 100)])
def test_getrf_base_kernel_variants_identical_pivots(sl, m, n, nb, panel, dist, variant, monkeypatch):
    _set_variant(monkeypatch, variant)
    monkeypatch.setenv("SB200_PANEL", panel)
    monkeypatch.setenv("SB200_GETRF_DIST", dist)
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, info_o = o.getrf(A0, nb, 32)
    assert info == info_o == 0
    assert piv == pivo, "pivot vectors differ from the oracle's"
    assert np.abs(A.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()


@pytest.mark.parametrize("variant", list(BASE_VARIANTS))
@pytest.mark.parametrize("panel", ["1", "2"])
def test_getrf_base_kernel_variants_zero_column_and_ties(sl, panel, variant, monkeypatch):
    _set_variant(monkeypatch, variant)
    monkeypatch.setenv("SB200_PANEL", panel)
    n, nb = 256, 64
    A0 = o.generate("rand", n, n, 3); A0[:, 100] = 0.0
    A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
    _, info = sl.getrf(A)
    assert info == o.getrf(A0, nb, 32)[2] == 101
    # equal magnitudes everywhere: the first (lowest) row must win every tie, as in the reference's scan order
    rng = np.random.default_rng(2)
    S = np.sign(rng.random((n, n)) - 0.5) + 2 * np.eye(n) * 0          # entries +-1
    B = sl.Matrix(n, n, nb); B.from_host(np.asfortranarray(S))
    piv, info = sl.getrf(B)
    LUo, pivo, info_o = o.getrf(S, nb, 32)
    assert info == info_o
    assert piv == pivo


def test_getrf_base_kernel_variants_nan_is_never_a_pivot_candidate(sl, monkeypatch):
    """a NaN below the diagonal is skipped by the max search (the reference's `abs > max` is false for it), a NaN ON the
    diagonal keeps the diagonal (every comparison with it is false): same pivots from every kernel"""
    n, nb = 256, 128
    A0 = o.generate("rand", n, n, 5)
    A0[200, 3] = np.nan
    pivs = []
    for variant in BASE_VARIANTS:
        _set_variant(monkeypatch, variant)
        A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
        piv, _ = sl.getrf(A)
        pivs.append(piv[0][:8])                                  # first panel, first columns
    assert all(p == pivs[0] for p in pivs)
    assert all(t * nb + off != 200 for (t, off) in pivs[0][:4])


def test_getrf_base_kernel_variants_tall_panel_many_ctas(sl, monkeypatch):
    """m_p = 8192 rows: 16 + 4 CTAs exchange through the tagged words; every variant gives the default's pivots and its
    factor up to rounding (the folded updates sum in a different order than the tile GEMM)"""
    n, nb = 8192, 512
    res = {}
    for variant in BASE_VARIANTS:
        _set_variant(monkeypatch, variant)
        A = sl.Matrix(n, n, nb).generate("rand", 42)
        piv, info = sl.getrf(A)
        assert info == 0
        res[variant] = (piv, A.to_host())
    p0, a0 = res["default"]
    for variant, (p1, a1) in res.items():
        assert p1 == p0, variant
        assert np.abs(a1 - a0).max() <= 1e-11 * np.abs(a0).max(), variant


@pytest.mark.parametrize("variant", ["default", "shared_memory_rows"])
def test_gesv_mixed_with_base_kernel_variants(sl, monkeypatch, variant):
    _set_variant(monkeypatch, variant)
    n, nb = 2048, 512
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    B = sl.Matrix(n, 10, nb).generate("rand", 43)
    X = sl.Matrix(n, 10, nb)
    info, it, piv, tm = sl.gesv_mixed(A, B, X)
    assert info == 0 and 0 <= it <= 30
    assert o.solve_residual(o.generate("rand", n, n, 42), X.to_host(), o.generate("rand", n, 10, 43)) <= 25 * EPS


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("diag", ["U", "N"])
@pytest.mark.parametrize("m,n", [(32, 32), (64, 64), (20, 100), (64, 300), (32, 1), (33, 512), (1, 7)])
def test_fused_row_trsm_small_triangle_direct_substitution(monkeypatch, t, diag, m, n):
    """SB200_TRSM_FUSED bit 2: na <= 64 (the U12 solves inside the recursive LU panel) by direct substitution."""
    from tests.gpu_util import DevTiles, fn, scal, stream, rng_tiles, NP, SC, c_int, c_i64, c_ptr
    monkeypatch.setenv("SB200_TRSM_FUSED", "4")
    rng = np.random.default_rng(4)
    batch = 2
    T = (rng.random((m, m)) / m + np.eye(m) * (1 + rng.random(m))).astype(NP[t])
    B = rng_tiles(rng, batch, m, n, t)
    alpha = 0.7
    ref = [o.trsm_tile("L", "L", "N", diag, alpha, T.astype(np.float64), b.astype(np.float64)) for b in B]
    dT, dB = DevTiles([T]), DevTiles(B)
    f = fn(f"sb200_trsm_batched_{t}", [c_int] * 5 + [c_i64, c_i64, SC[t], c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr])
    assert f(ord("C"), ord("L"), ord("L"), ord("N"), ord(diag), m, n, scal(t, alpha), dT.t[0].data_ptr(), m,
             dB.p, m, batch, None, stream()) == 0
    eps = EPS if t == "d" else float(np.finfo(np.float32).eps)
    for x, r in zip(dB.get(), ref):
        assert np.abs(x - r).max() <= 200 * eps * np.abs(r).max()


@pytest.mark.parametrize("panel", ["1", "2"])
@pytest.mark.parametrize("m,n,nb", [(1024, 1024, 256), (2048, 2048, 512), (700, 300, 128), (1100, 1100, 512)])
def test_getrf_with_all_row_solve_candidates_identical_pivots(sl, monkeypatch, m, n, nb, panel):
    monkeypatch.setenv("SB200_TRSM_FUSED", "6")
    monkeypatch.setenv("SB200_PANEL", panel)
    A = sl.Matrix(m, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    A0 = o.generate("rand", m, n, 42)
    LUo, pivo, info_o = o.getrf(A0, nb, 32)
    assert info == info_o == 0
    assert piv == pivo, "pivot vectors differ from the oracle's"
    assert np.abs(A.to_host() - LUo).max() <= GETRF_TOL * np.abs(LUo).max()


@pytest.mark.parametrize("t", ["s", "c", "z"])
@pytest.mark.parametrize("layout", ["C", "R"])
def test_permute_rows_other_types(t, layout):
    """sb200_permute_rows_{s,c,z}: the same launch_laswp template as the validated _d entry (internal_swap.cc:674-688)."""
    import torch
    from tests.gpu_util import DevTiles, fn, stream, NP, c_int, c_i64, c_ptr
    rng = np.random.default_rng(6)
    mt, ncb, mb, nc = 3, 2, 32, 24
    M = rng.random((mt * mb, ncb * nc))
    if t in "cz":
        M = M + 1j * rng.random(M.shape)
    M = M.astype(NP[t])
    npiv = mb
    piv = [(int(rng.integers(0, mt)), int(rng.integers(0, mb))) for _ in range(npiv)]
    piv = [(ti, off) if ti * mb + off >= j else (0, j) for j, (ti, off) in enumerate(piv)]
    ref = M.copy()
    for j in range(npiv):
        r2 = piv[j][0] * mb + piv[j][1]
        ref[[j, r2]] = ref[[r2, j]]
    tiles = []
    for jb in range(ncb):
        for tb in range(mt):
            blk = M[tb * mb:(tb + 1) * mb, jb * nc:(jb + 1) * nc]
            tiles.append(np.asfortranarray(blk if layout == "C" else blk.T))
    d = DevTiles(tiles)
    pt = torch.tensor([p[0] for p in piv], dtype=torch.int64, device="cuda")
    po = torch.tensor([p[1] for p in piv], dtype=torch.int64, device="cuda")
    f = fn(f"sb200_permute_rows_{t}", [c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr])
    ld = mb if layout == "C" else nc
    assert f(ord(layout), 1, npiv, pt.data_ptr(), po.data_ptr(), d.p, mt, ncb, mb, nc, ld, stream()) == 0
    out = d.get()
    got = np.zeros_like(M)
    for jb in range(ncb):
        for tb in range(mt):
            blk = out[tb + jb * mt]
            got[tb * mb:(tb + 1) * mb, jb * nc:(jb + 1) * nc] = blk if layout == "C" else blk.T
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("rsqrt", ["1", "0"])
@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("n", [37, 64, 200, 512, 130])
def test_potrf_tile_diag_block_forms(monkeypatch, rsqrt, t, n):
    """the 64 x 64 diagonal-block Cholesky with one rsqrt per column (default) and with division + sqrt"""
    monkeypatch.setenv("SB200_DIAG_RSQRT", rsqrt)
    rng = np.random.default_rng(5)
    G = rng.random((n, n))
    A = (G @ G.T + n * np.eye(n)).astype(np.float64 if t == "d" else np.float32)
    out, info = _potrf_tile(A, n, n, t)
    assert info == 0
    ref = np.linalg.cholesky(A.astype(np.float64))
    eps = EPS if t == "d" else float(np.finfo(np.float32).eps)
    assert np.abs(np.tril(out) - ref).max() <= 50 * eps * np.abs(ref).max()
    assert np.array_equal(np.triu(out, 1), np.triu(A, 1))


@pytest.mark.parametrize("rsqrt", ["1", "0"])
def test_potrf_tile_diag_block_forms_info(monkeypatch, rsqrt):
    monkeypatch.setenv("SB200_DIAG_RSQRT", rsqrt)
    A = np.eye(128); A[70, 70] = -1.0
    _, info = _potrf_tile(A, 128)
    assert info == 71


@pytest.mark.parametrize("m,n,k,nb", [(1024, 1024, 1024, 256), (700, 900, 500, 128), (2048, 1536, 1024, 512), (300, 200, 100, 512)])
def test_gemm_transposed_b_panel_is_bitwise_the_default(sl, monkeypatch, m, n, k, nb):
    import torch

    def run():
        A = sl.Matrix(m, k, nb).generate("rand", 1)
        B = sl.Matrix(k, n, nb).generate("rand", 2)
        C = sl.Matrix(m, n, nb).generate("rand", 3)
        sl.gemm(3.1, A, B, 2.7, C)
        return C.to_host()

    monkeypatch.delenv("SB200_GEMM_BT", raising=False)
    c0 = run()
    monkeypatch.setenv("SB200_GEMM_BT", "1")
    c1 = run()
    assert np.array_equal(c0, c1)
    a, b, c = (o.generate("rand", *shape, seed) for shape, seed in (((m, k), 1), ((k, n), 2), ((m, n), 3)))
    assert o.gemm_check(3.1, a, b, 2.7, c, c1) <= 3 * EPS


@pytest.mark.parametrize("dist", ["0", "1"])
@pytest.mark.parametrize("m,n,nb", [(2048, 2048, 512), (1100, 1100, 256), (700, 1000, 128), (1000, 700, 128)])
def test_getrf_transposed_u_row_is_bitwise_the_default(sl, monkeypatch, m, n, nb, dist):
    monkeypatch.setenv("SB200_GETRF_DIST", dist)            # 1 = the p x q driver (getrf_dist.cu) on one rank

    def run():
        A = sl.Matrix(m, n, nb).generate("rand", 42)
        piv, info = sl.getrf(A)
        return piv, info, A.to_host()

    monkeypatch.delenv("SB200_GEMM_BT", raising=False)
    p0, i0, a0 = run()
    monkeypatch.setenv("SB200_GEMM_BT", "1")
    p1, i1, a1 = run()
    assert i0 == i1 == 0 and p0 == p1
    assert np.array_equal(a0, a1)


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
def test_her2k_matches_reference_golden_and_oracle(sl, golden_dir, t):
    """golden her2k_{d,z}.npz were written by the unmodified reference (slate::her2k, HostTask); s / c against the oracle"""
    from tests.gpu_util import NP
    n, k, nb = 200, 100, 64
    al = (3.141592653589793 + 1.414213562373095j) if t in "cz" else 3.141592653589793
    be = 2.718281828459045
    A = sl.Matrix(n, k, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, k, nb, dtype=t).generate("rand", 43)
    C = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 44)
    sl.her2k(al, A, B, be, C)
    out = np.tril(C.to_host())
    a, b, c = (o.generate("rand", *shape, seed, NP[t]) for shape, seed in (((n, k), 42), ((n, k), 43), ((n, n), 44)))
    wide = np.complex128 if t in "cz" else np.float64
    ref = np.tril(o.her2k(al, a.astype(wide), b.astype(wide), be, np.tril(c).astype(wide), nb))
    eps = EPS if t in "dz" else float(np.finfo(np.float32).eps)
    assert np.abs(out - ref).max() <= 64 * eps * np.abs(ref).max()
    if t in "dz":
        g = np.load(os.path.join(golden_dir, f"her2k_{t}.npz"))
        assert np.abs(out - np.tril(g["out"])).max() <= 64 * EPS * np.abs(g["out"]).max()
    if t in "cz":
        assert np.all(np.diag(out).imag == 0)


def test_her2k_larger_and_ragged(sl):
    n, k, nb = 1100, 700, 256
    A = sl.Matrix(n, k, nb).generate("rand", 1)
    B = sl.Matrix(n, k, nb).generate("rand", 2)
    C = sl.HermitianMatrix(n, nb).generate("rand", 3)
    sl.her2k(0.5, A, B, 1.5, C)
    a, b, c = o.generate("rand", n, k, 1), o.generate("rand", n, k, 2), np.tril(o.generate("rand", n, n, 3))
    ref = np.tril(0.5 * (a @ b.T) + 0.5 * (b @ a.T) + 1.5 * (c + np.tril(c, -1).T))
    assert np.abs(np.tril(C.to_host()) - ref).max() <= 3 * np.sqrt(2 * k) * EPS * 4 * np.abs(ref).max()


@pytest.mark.parametrize("on_device", [False, True])
@pytest.mark.parametrize("kind,m,n,nb", [("G", 300, 200, 64), ("G", 512, 512, 128), ("H", 300, 300, 64)])
def test_scalapack_local_array_round_trip(sl, kind, m, n, nb, on_device):
    """1 x 1 grid: the local array IS the global matrix (column-major, lld > m); gather into tiles, compare with from_host,
    scatter back: bit-exact data movement (Matrix::fromScaLAPACK, include/slate/Matrix.hh:75-99)."""
    import torch
    rng = np.random.default_rng(8)
    G = rng.random((m, n))
    lld = m + 5
    loc = torch.zeros((n, lld), dtype=torch.float64)                  # [local columns][lld] == column-major lld x n
    loc[:, :m] = torch.from_numpy(np.ascontiguousarray(G.T))
    if on_device:
        loc = loc.cuda()
    A = sl.HermitianMatrix(n, nb) if kind == "H" else sl.Matrix(m, n, nb)
    A.from_scalapack(loc, lld)
    got = A.to_host()
    B = sl.HermitianMatrix(n, nb) if kind == "H" else sl.Matrix(m, n, nb)
    B.from_host(np.asfortranarray(G))
    assert np.array_equal(got, B.to_host())
    back = torch.full((n, lld), -1.0, dtype=torch.float64)
    if on_device:
        back = back.cuda()
    A.to_scalapack(back, lld)
    bk = back.cpu().numpy()[:, :m].T
    if kind == "H":
        assert np.array_equal(np.tril(bk), np.tril(G))                # only the stored (lower) tiles come back
    else:
        assert np.array_equal(bk, G)
    assert np.all(back.cpu().numpy()[:, m:] == -1.0)                 # the padding rows of the local array are untouched


@pytest.mark.parametrize("dist", ["0", "1"])
@pytest.mark.parametrize("n,nb", [(300, 128), (1024, 256), (2048, 512)])
def test_getrf_nopiv_matches_reference_golden_and_oracle(sl, golden_dir, monkeypatch, n, nb, dist):
    monkeypatch.setenv("SB200_GETRF_DIST", dist)
    A = sl.Matrix(n, n, nb).generate("rand_dominant", 42)
    assert sl.getrf_nopiv(A) == 0
    LU = A.to_host()
    A0 = o.generate("rand_dominant", n, n, 42)
    LUo, info = o.getrf_nopiv(A0, nb)
    assert info == 0
    assert np.abs(LU - LUo).max() <= 64 * EPS * np.abs(LUo).max()
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    assert np.abs(L @ U - A0).max() <= 64 * EPS * np.abs(A0).max()
    if (n, nb) == (300, 128):
        g = np.load(os.path.join(golden_dir, "getrf_nopiv_d.npz"))          # written by the unmodified reference
        assert np.abs(LU - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


def test_getrf_nopiv_zero_pivot_info_and_pivoting_is_back_afterwards(sl):
    n, nb = 256, 64
    Z = o.generate("rand_dominant", n, n, 3); Z[:, 100] = 0.0
    A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(Z))
    assert sl.getrf_nopiv(A) == o.getrf_nopiv(Z, nb)[1] == 101
    B = sl.Matrix(n, n, nb).generate("rand", 42)                             # the switch is per call: getrf pivots again
    piv, info = sl.getrf(B)
    assert info == 0 and piv == o.getrf(o.generate("rand", n, n, 42), nb, 32)[1]


@pytest.mark.parametrize("t", ["z", "c", "d"])
@pytest.mark.parametrize("routine", ["syrk", "syr2k"])
def test_symmetric_rank_updates_match_reference_golden_and_oracle(sl, golden_dir, routine, t):
    """complex-symmetric syrk / syr2k (no conjugation, complex diagonal kept); golden {syrk,syr2k}_z.npz from the reference"""
    from tests.gpu_util import NP
    n, k, nb = 200, 100, 64
    al = (3.141592653589793 + 1.414213562373095j) if t in "cz" else 3.141592653589793
    be = (2.718281828459045 + 1.732050807568877j) if t in "cz" else 2.718281828459045
    A = sl.Matrix(n, k, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, k, nb, dtype=t).generate("rand", 43)
    C = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 44)
    if routine == "syrk":
        sl.syrk(al, A, be, C)
    else:
        sl.syr2k(al, A, B, be, C)
    out = np.tril(C.to_host())
    wide = np.complex128 if t in "cz" else np.float64
    a, b, c = (o.generate("rand", *shape, seed, NP[t]).astype(wide) for shape, seed in (((n, k), 42), ((n, k), 43), ((n, n), 44)))
    ref = np.tril(o.syrk(al, a, be, np.tril(c), nb) if routine == "syrk" else o.syr2k(al, a, b, be, np.tril(c), nb))
    eps = EPS if t in "dz" else float(np.finfo(np.float32).eps)
    assert np.abs(out - ref).max() <= 64 * eps * np.abs(ref).max()
    if t == "z":
        g = np.load(os.path.join(golden_dir, f"{routine}_z.npz"))
        assert np.abs(out - np.tril(g["out"])).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("t", ["z", "c", "d"])
def test_symm_matches_reference_golden_and_oracle(sl, golden_dir, t):
    from tests.gpu_util import NP
    n, nb, nrhs = 192, 64, 70
    al = (3.141592653589793 + 1.414213562373095j) if t in "cz" else 3.141592653589793
    be = (2.718281828459045 + 1.732050807568877j) if t in "cz" else 2.718281828459045
    A = sl.HermitianMatrix(n, nb, dtype=t).generate("rand", 42)
    B = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 43)
    C = sl.Matrix(n, nrhs, nb, dtype=t).generate("rand", 44)
    sl.symm(al, A, B, be, C)
    wide = np.complex128 if t in "cz" else np.float64
    a = np.tril(o.generate("rand", n, n, 42, NP[t])).astype(wide)
    b, c = (o.generate("rand", n, nrhs, seed, NP[t]).astype(wide) for seed in (43, 44))
    ref = o.symm(al, a, b, be, c, nb)
    eps = EPS if t in "dz" else float(np.finfo(np.float32).eps)
    out = C.to_host()
    assert np.abs(out - ref).max() <= 64 * eps * np.abs(ref).max()
    if t == "z":
        g = np.load(os.path.join(golden_dir, "symm_z.npz"))
        assert np.abs(out - g["out"]).max() <= 64 * EPS * np.abs(g["out"]).max()


@pytest.mark.parametrize("t", ["d", "z", "s"])
@pytest.mark.parametrize("diag", ["N", "U"])
@pytest.mark.parametrize("m,n,nb", [(200, 70, 64), (512, 512, 128), (300, 1000, 256)])
def test_trmm_left_lower_matches_oracle_and_reference_golden(sl, golden_dir, t, diag, m, n, nb):
    from tests.gpu_util import NP
    al = (3.141592653589793 + 1.414213562373095j) if t == "z" else 3.141592653589793
    A = sl.HermitianMatrix(m, nb, dtype=t).generate("rand", 42)             # its lower tiles are the triangle
    B = sl.Matrix(m, n, nb, dtype=t).generate("rand", 43)
    sl.trmm(al, A, B, diag=diag)
    wide = np.complex128 if t == "z" else np.float64
    a = np.tril(o.generate("rand", m, m, 42, NP[t])).astype(wide)
    b = o.generate("rand", m, n, 43, NP[t]).astype(wide)
    ref = o.trmm(al, a, b, nb, unit=(diag == "U"))
    eps = EPS if t in "dz" else float(np.finfo(np.float32).eps)
    out = B.to_host()
    assert np.abs(out - ref).max() <= 3 * np.sqrt(m) * eps * 4 * np.abs(ref).max()
    if (t, diag, m, n, nb) == ("d", "N", 200, 70, 64):
        g = np.load(os.path.join(golden_dir, "trmm_d.npz"))                  # written by the unmodified reference
        assert np.abs(out - g["out"]).max() <= 3 * np.sqrt(m) * EPS * 4 * np.abs(g["out"]).max()


def test_trmm_unsupported_variants_say_so(sl):
    A = sl.HermitianMatrix(64, 32); B = sl.Matrix(64, 8, 32)
    with pytest.raises(sl.SB200Error):
        sl.trmm(1.0, A, B, side="R")                       # Side::Right needs A.n == B.n (8 != 64): SB200_EINVAL
    with pytest.raises(sl.SB200Error):
        sl.trmm(1.0, A, B, uplo="U")                       # lower storage only: SB200_ENOTSUP
