"""GPU parity tests (pytest -m gpu): every C-ABI kernel against the numpy oracle on the same
seeded inputs, through the C ABI (ctypes -> libslate_b200.so).  Shapes follow the reference's
unit tests (unit_test/test_internal_blas.cc:283-452: m=80, n=64, k=16, tol 3 sqrt(k) eps;
test_geadd.cc / test_gescale.cc / test_geset.cc / test_gecopy.cc / test_norm.cc) plus ragged,
odd and unaligned cases and the production tile sizes."""
import numpy as np
import pytest

from oracle import slate_oracle as o
from tests.gpu_util import (DevTiles, fn, scal, sync, stream, rng_tiles, dev_zeros, NP, REAL, SC, RSC,
                            c_i64, c_int, c_dbl, c_flt, c_ptr)

pytestmark = pytest.mark.gpu
EPS = {"s": np.finfo(np.float32).eps, "d": np.finfo(np.float64).eps,
       "c": np.finfo(np.float32).eps, "z": np.finfo(np.float64).eps}


def opmat(a, op):
    return a if op == "N" else (a.T if op == "T" else a.conj().T)


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
@pytest.mark.parametrize("layout", ["C", "R"])
@pytest.mark.parametrize("opA,opB", [("N", "N"), ("N", "T"), ("T", "N"), ("C", "C"), ("N", "C")])
@pytest.mark.parametrize("m,n,k,batch", [(80, 64, 16, 3), (130, 70, 36, 2), (77, 53, 19, 2), (256, 256, 256, 2), (0, 5, 3, 1)])
def test_gemm_batched(t, layout, opA, opB, m, n, k, batch):
    rng = np.random.default_rng(1)
    shpA = (m, k) if opA == "N" else (k, m)
    shpB = (k, n) if opB == "N" else (n, k)
    A = rng_tiles(rng, batch, *shpA, t); B = rng_tiles(rng, batch, *shpB, t); C = rng_tiles(rng, batch, m, n, t)
    alpha, beta = (3.1 + 1.4j, 2.7 + 1.7j) if t in "cz" else (3.1, 2.7)
    ref = [alpha * (opmat(a, opA) @ opmat(b, opB)) + beta * c for a, b, c in zip(A, B, C)]
    # row-major storage of X == column-major storage of X^T
    st = (lambda x: x) if layout == "C" else (lambda x: np.asfortranarray(x.T))
    dA, dB, dC = DevTiles([st(a) for a in A]), DevTiles([st(b) for b in B]), DevTiles([st(c) for c in C])
    ld = lambda shp: max(1, shp[0] if layout == "C" else shp[1])
    f = fn(f"sb200_gemm_batched_{t}", [c_int] * 3 + [c_i64] * 3 + [SC[t], c_ptr, c_i64, c_ptr, c_i64, SC[t], c_ptr, c_i64, c_i64, c_ptr])
    rc = f(ord(layout), ord(opA), ord(opB), m, n, k, scal(t, alpha), dA.p, ld(shpA), dB.p, ld(shpB),
           scal(t, beta), dC.p, ld((m, n)), batch, stream())
    assert rc == 0
    if m == 0:
        return
    out = dC.get()
    if layout == "R":
        out = [x.T for x in out]
    tol = 3 * np.sqrt(max(k, 1)) * EPS[t] * 4
    for x, r in zip(out, ref):
        assert np.abs(x - r).max() <= tol * np.abs(r).max()


def test_gemm_unaligned_pointers_and_odd_ld():
    """Producer fallback path: 8-byte aligned pointers, odd leading dimensions."""
    import torch
    rng = np.random.default_rng(2)
    m, n, k, lda, ldb, ldc = 67, 45, 23, 69, 25, 71
    buf = torch.zeros(3 * 80 * 80 + 3, dtype=torch.float64, device="cuda")
    A = rng.random((lda, k)); B = rng.random((ldb, n)); C = rng.random((ldc, n))
    offs = [1, 80 * 80 + 2, 2 * 80 * 80 + 3]          # odd element offsets -> 8-byte alignment only
    for off, X in zip(offs, (A, B, C)):
        buf[off:off + X.size] = torch.from_numpy(np.asfortranarray(X).ravel(order="F")).cuda()
    ptr = lambda i: torch.tensor([buf.data_ptr() + 8 * offs[i]], dtype=torch.int64, device="cuda")
    pA, pB, pC = ptr(0), ptr(1), ptr(2)
    f = fn("sb200_gemm_batched_d", [c_int] * 3 + [c_i64] * 3 + [c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord("C"), ord("N"), ord("N"), m, n, k, -1.0, pA.data_ptr(), lda, pB.data_ptr(), ldb, 0.5, pC.data_ptr(), ldc, 1, stream()) == 0
    sync()
    out = buf[offs[2]:offs[2] + ldc * n].cpu().numpy().reshape(ldc, n, order="F")
    ref = -A[:m] @ B[:k] + 0.5 * C[:m]
    assert np.abs(out[:m] - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(out[m:], C[m:])            # rows beyond m untouched


@pytest.mark.parametrize("t", ["d", "z", "s"])
@pytest.mark.parametrize("layout", ["C", "R"])
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("op", ["N", "C"])
@pytest.mark.parametrize("n,k", [(64, 16), (150, 40), (512, 512)])
def test_herk_batched(t, layout, uplo, op, n, k):
    rng = np.random.default_rng(3)
    batch = 2
    shpA = (n, k) if op == "N" else (k, n)
    A = rng_tiles(rng, batch, *shpA, t); C = rng_tiles(rng, batch, n, n, t)
    alpha, beta = -1.0, 1.0
    opc = "C" if t == "z" else "T"
    full = [alpha * (opmat(a, "N" if op == "N" else opc) @ opmat(a, opc if op == "N" else "N")) + beta * c for a, c in zip(A, C)]
    mask = np.tril(np.ones((n, n), bool)) if uplo == "L" else np.triu(np.ones((n, n), bool))
    st = (lambda x: x) if layout == "C" else (lambda x: np.asfortranarray(x.T))
    dA, dC = DevTiles([st(a) for a in A]), DevTiles([st(c) for c in C])
    ld = lambda shp: shp[0] if layout == "C" else shp[1]
    f = fn(f"sb200_herk_batched_{t}", [c_int] * 3 + [c_i64] * 2 + [RSC[t], c_ptr, c_i64, RSC[t], c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord(layout), ord(uplo), ord(op), n, k, RSC[t](alpha), dA.p, ld(shpA), RSC[t](beta), dC.p, n, batch, stream()) == 0
    out = dC.get()
    if layout == "R":
        out = [x.T for x in out]
    tol = 3 * np.sqrt(k) * EPS[t] * 4
    for x, r, c in zip(out, full, C):
        exp = np.where(mask, r, c)                   # other triangle untouched
        if t == "z":
            d = np.arange(n); exp[d, d] = exp[d, d].real
        assert np.abs(x - exp).max() <= tol * np.abs(r).max()


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("layout", ["C", "R"])
@pytest.mark.parametrize("side", ["L", "R"])
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("op", ["N", "T"])
@pytest.mark.parametrize("diag", ["N", "U"])
@pytest.mark.parametrize("m,n", [(64, 48), (200, 136), (77, 53), (512, 512)])
def test_trsm_batched(t, layout, side, uplo, op, diag, m, n):
    rng = np.random.default_rng(4)
    batch = 2
    na = m if side == "L" else n
    T = (rng.random((na, na)) / na + np.eye(na) * (1 + rng.random(na))).astype(NP[t])
    B = rng_tiles(rng, batch, m, n, t)
    alpha = 0.7
    ref = [o.trsm_tile(side, uplo, op, diag, alpha, T.astype(np.float64), b.astype(np.float64)) for b in B]
    st = (lambda x: x) if layout == "C" else (lambda x: np.asfortranarray(x.T))
    dT, dB = DevTiles([st(T)]), DevTiles([st(b) for b in B])
    f = fn(f"sb200_trsm_batched_{t}", [c_int] * 5 + [c_i64, c_i64, SC[t], c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr])
    assert f(ord(layout), ord(side), ord(uplo), ord(op), ord(diag), m, n, scal(t, alpha), dT.t[0].data_ptr(), na,
             dB.p, m if layout == "C" else n, batch, None, stream()) == 0
    out = dB.get()
    if layout == "R":
        out = [x.T for x in out]
    for x, r in zip(out, ref):
        assert np.abs(x - r).max() <= 200 * EPS[t] * np.abs(r).max()


@pytest.mark.parametrize("t", ["d", "s"])
@pytest.mark.parametrize("n", [37, 64, 200, 512])
def test_potrf_tile(t, n):
    rng = np.random.default_rng(5)
    G = rng.random((n, n))
    A = (G @ G.T + n * np.eye(n)).astype(NP[t])
    dA = DevTiles([A.copy()])
    info = dev_zeros(1, np.int32)
    f = fn(f"sb200_potrf_tile_{t}", [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr])
    assert f(ord("L"), n, dA.t[0].data_ptr(), n, info.data_ptr(), None, stream()) == 0
    out = dA.get()[0]
    assert int(info.cpu()[0]) == 0
    ref = np.linalg.cholesky(A.astype(np.float64))
    assert np.abs(np.tril(out) - ref).max() <= 50 * EPS[t] * np.abs(ref).max()
    assert np.array_equal(np.triu(out, 1), np.triu(A, 1))         # strict upper triangle untouched


def test_potrf_tile_reports_first_bad_minor():
    A = np.eye(128); A[70, 70] = -1.0
    dA = DevTiles([A]); info = dev_zeros(1, np.int32)
    f = fn("sb200_potrf_tile_d", [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr])
    assert f(ord("L"), 128, dA.t[0].data_ptr(), 128, info.data_ptr(), None, stream()) == 0
    sync()
    assert int(info.cpu()[0]) == 71


@pytest.mark.parametrize("layout", ["C", "R"])
@pytest.mark.parametrize("forward", [1, 0])
def test_permute_rows(layout, forward):
    """Fused row interchanges == the reference's sequential swaps (internal_swap.cc:674-688)."""
    import torch
    rng = np.random.default_rng(6)
    mt, ncb, mb, nc = 3, 2, 32, 24
    M = rng.random((mt * mb, ncb * nc))
    npiv = mb
    piv = [(int(rng.integers(0, mt)), int(rng.integers(0, mb))) for _ in range(npiv)]
    piv = [(ti, off) if ti * mb + off >= j else (0, j) for j, (ti, off) in enumerate(piv)]
    ref = M.copy()
    order = range(npiv) if forward else range(npiv - 1, -1, -1)
    for j in order:
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
    f = fn("sb200_permute_rows_d", [c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr])
    ld = mb if layout == "C" else nc
    assert f(ord(layout), forward, npiv, pt.data_ptr(), po.data_ptr(), d.p, mt, ncb, mb, nc, ld, stream()) == 0
    out = d.get()
    got = np.zeros_like(M)
    for jb in range(ncb):
        for tb in range(mt):
            blk = out[tb + jb * mt]
            got[tb * mb:(tb + 1) * mb, jb * nc:(jb + 1) * nc] = blk if layout == "C" else blk.T
    assert np.array_equal(got, ref)                  # pure data movement: bit-exact


# ------------------------------------------------------------------------------- tile kernels
@pytest.mark.parametrize("t", ["d", "s", "z"])
@pytest.mark.parametrize("m,n", [(64, 48), (257, 129), (512, 512), (1, 7)])
def test_geadd_gescale_geset(t, m, n):
    rng = np.random.default_rng(7)
    batch = 3
    A = rng_tiles(rng, batch, m, n, t); B = rng_tiles(rng, batch, m, n, t)
    alpha, beta = (1.5 - 0.5j, 0.25 + 2j) if t == "z" else (1.5, 0.25)
    dA, dB = DevTiles(A), DevTiles(B)
    f = fn(f"sb200_geadd_batched_{t}", [c_i64, c_i64, SC[t], c_ptr, c_i64, SC[t], c_ptr, c_i64, c_i64, c_ptr])
    assert f(m, n, scal(t, alpha), dA.p, m, scal(t, beta), dB.p, m, batch, stream()) == 0
    for x, a, b in zip(dB.get(), A, B):
        r = o.geadd(alpha, a, beta, b)
        assert np.abs(x - r).max() <= 4 * EPS[t] * np.abs(r).max()
    f = fn(f"sb200_gescale_batched_{t}", [c_i64, c_i64, SC[t], SC[t], c_ptr, c_i64, c_i64, c_ptr])
    assert f(m, n, scal(t, 3.0), scal(t, 2.0), dA.p, m, batch, stream()) == 0
    for x, a in zip(dA.get(), A):
        assert np.abs(x - o.gescale(3.0, 2.0, a)).max() <= 4 * EPS[t] * np.abs(a).max() * 1.5
    f = fn(f"sb200_geset_batched_{t}", [c_i64, c_i64, SC[t], SC[t], c_ptr, c_i64, c_i64, c_ptr])
    assert f(m, n, scal(t, 0.5), scal(t, 9.0), dA.p, m, batch, stream()) == 0
    for x in dA.get():
        assert np.array_equal(x, o.geset(0.5, 9.0, m, n, NP[t]))


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("m,n", [(64, 64), (100, 70), (70, 100)])
def test_trapezoid_kernels(uplo, m, n):
    rng = np.random.default_rng(8)
    batch = 2
    A = rng_tiles(rng, batch, m, n, "d"); B = rng_tiles(rng, batch, m, n, "d")
    dA, dB = DevTiles(A), DevTiles(B)
    f = fn("sb200_tzadd_batched_d", [c_int, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord(uplo), m, n, 2.0, dA.p, m, -1.0, dB.p, m, batch, stream()) == 0
    for x, a, b in zip(dB.get(), A, B):
        assert np.allclose(x, o.tzadd(uplo, 2.0, a, -1.0, b), rtol=1e-15, atol=0)
    f = fn("sb200_tzscale_batched_d", [c_int, c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord(uplo), m, n, 1.0, 4.0, dA.p, m, batch, stream()) == 0
    for x, a in zip(dA.get(), A):
        assert np.array_equal(x, o.tzscale(uplo, 1.0, 4.0, a))
    f = fn("sb200_tzset_batched_d", [c_int, c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord(uplo), m, n, 0.0, 1.0, dA.p, m, batch, stream()) == 0
    for x, a in zip(dA.get(), [o.tzscale(uplo, 1.0, 4.0, a) for a in A]):
        assert np.array_equal(x, o.tzset(uplo, 0.0, 1.0, a))
    C = rng_tiles(rng, batch, m, n, "d")
    dC = DevTiles(C); dS = DevTiles([np.zeros((m, n), np.float32, order="F") for _ in range(batch)])
    f = fn("sb200_tzcopy_batched_ds", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord(uplo), m, n, dC.p, m, dS.p, m, batch, stream()) == 0
    for x, c in zip(dS.get(), C):
        assert np.array_equal(x, o.tzcopy(uplo, c, np.zeros((m, n), np.float32), np.float32))


@pytest.mark.parametrize("pair", ["dd", "ds", "sd", "ss", "zz", "zc", "cz"])
def test_gecopy_converting(pair):
    rng = np.random.default_rng(9)
    m, n, batch = 130, 70, 2
    A = rng_tiles(rng, batch, m, n, pair[0])
    dA = DevTiles(A); dB = DevTiles([np.zeros((m, n), NP[pair[1]], order="F") for _ in range(batch)])
    f = fn(f"sb200_gecopy_batched_{pair}", [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    assert f(m, n, dA.p, m, dB.p, m, batch, stream()) == 0
    for x, a in zip(dB.get(), A):
        assert np.array_equal(x, a.astype(NP[pair[1]]))          # IEEE conversion: bit-exact


def test_gescale_row_col():
    rng = np.random.default_rng(10)
    m, n, batch = 96, 80, 2
    A = rng_tiles(rng, batch, m, n, "d")
    R = [rng.random((m, 1)) for _ in range(batch)]; Cc = [rng.random((n, 1)) for _ in range(batch)]
    f = fn("sb200_gescale_row_col_batched_d", [c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr])
    for eq in "RCB":
        dA, dR, dC = DevTiles(A), DevTiles(R), DevTiles(Cc)
        assert f(ord(eq), m, n, dR.p, dC.p, dA.p, m, batch, stream()) == 0
        for x, a, r, c in zip(dA.get(), A, R, Cc):
            assert np.allclose(x, o.gescale_row_col(eq, r[:, 0], c[:, 0], a), rtol=2e-16, atol=0)


@pytest.mark.parametrize("t", ["d", "s", "z"])
@pytest.mark.parametrize("m,n", [(64, 64), (100, 36), (33, 257), (512, 512)])
def test_transposes(t, m, n):
    rng = np.random.default_rng(11)
    batch = 2
    A = rng_tiles(rng, batch, m, n, t)
    dA = DevTiles(A); dT = DevTiles([np.zeros((n, m), NP[t], order="F") for _ in range(batch)])
    if t == "z":
        f = fn("sb200_transpose_batched_z", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
        assert f(1, m, n, dA.p, m, dT.p, n, batch, stream()) == 0
        exp = [a.conj().T for a in A]
    else:
        f = fn(f"sb200_transpose_batched_{t}", [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
        assert f(m, n, dA.p, m, dT.p, n, batch, stream()) == 0
        exp = [a.T for a in A]
    for x, e in zip(dT.get(), exp):
        assert np.array_equal(x, e)
    if m == n:
        if t == "z":
            f = fn("sb200_transpose_inplace_batched_z", [c_int, c_i64, c_ptr, c_i64, c_i64, c_ptr])
            assert f(0, n, dA.p, n, batch, stream()) == 0
        else:
            f = fn(f"sb200_transpose_inplace_batched_{t}", [c_i64, c_ptr, c_i64, c_i64, c_ptr])
            assert f(n, dA.p, n, batch, stream()) == 0
        for x, a in zip(dA.get(), A):
            assert np.array_equal(x, a.T)


@pytest.mark.parametrize("t", ["d", "s", "z"])
@pytest.mark.parametrize("m,n", [(64, 48), (200, 136), (512, 512)])
def test_genorm(t, m, n):
    rng = np.random.default_rng(12)
    batch = 3
    A = rng_tiles(rng, batch, m, n, t)
    A[1][m // 2, n // 3] = -7.5
    dA = DevTiles(A)
    rt = REAL[t]
    f = fn(f"sb200_genorm_batched_{t}", [c_int, c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    tol = 64 * EPS[t]
    for norm, ldv in (("M", 1), ("O", n), ("I", m), ("F", 2)):
        vals = dev_zeros(batch * ldv, rt)
        assert f(ord(norm), ord("M"), m, n, dA.p, m, vals.data_ptr(), ldv, batch, stream()) == 0
        sync()
        v = vals.cpu().numpy().reshape(batch, ldv)
        for k, a in enumerate(A):
            r = o.genorm(norm, a.astype(np.complex128 if t == "z" else np.float64))
            if norm == "F":
                assert abs(v[k, 0] * np.sqrt(v[k, 1]) - r[0] * np.sqrt(r[1])) <= tol * r[0] * np.sqrt(r[1])
            elif norm == "M":
                assert abs(v[k, 0] - r) <= 2 * EPS[t] * r
            else:
                assert np.abs(v[k] - r).max() <= tol * np.abs(r).max()
    vals = dev_zeros(batch * n, rt)
    assert f(ord("M"), ord("C"), m, n, dA.p, m, vals.data_ptr(), n, batch, stream()) == 0
    sync()
    v = vals.cpu().numpy().reshape(batch, n)
    for k, a in enumerate(A):
        assert np.abs(v[k] - o.genorm_colmax(a)).max() <= 2 * EPS[t] * np.abs(a).max()


def test_genorm_max_propagates_nan():
    A = [np.asfortranarray(np.ones((40, 30))), np.asfortranarray(np.ones((40, 30)))]
    A[1][7, 9] = np.nan
    dA = DevTiles(A); vals = dev_zeros(2, np.float64)
    f = fn("sb200_genorm_batched_d", [c_int, c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord("M"), ord("M"), 40, 30, dA.p, 40, vals.data_ptr(), 1, 2, stream()) == 0
    sync()
    v = vals.cpu().numpy()
    assert v[0] == 1.0 and np.isnan(v[1])


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [64, 150])
def test_structured_norms(uplo, n):
    rng = np.random.default_rng(13)
    batch = 2
    A = rng_tiles(rng, batch, n, n, "d")
    Z = rng_tiles(rng, batch, n, n, "z")
    dA, dZ = DevTiles(A), DevTiles(Z)
    tol = 64 * EPS["d"]
    he = fn("sb200_henorm_batched_z", [c_int, c_int, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    sy = fn("sb200_synorm_batched_d", [c_int, c_int, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    tr = fn("sb200_trnorm_batched_d", [c_int, c_int, c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    for norm, ldv in (("M", 1), ("O", n), ("F", 2)):
        for fcall, tiles, dev in ((he, Z, dZ), (sy, A, dA)):
            vals = dev_zeros(batch * ldv, np.float64)
            assert fcall(ord(norm), ord(uplo), n, dev.p, n, vals.data_ptr(), ldv, batch, stream()) == 0
            sync()
            v = vals.cpu().numpy().reshape(batch, ldv)
            for k, a in enumerate(tiles):
                r = o.henorm(norm, uplo, a)
                if norm == "F":
                    assert abs(v[k, 0] * np.sqrt(v[k, 1]) - r[0] * np.sqrt(r[1])) <= tol * r[0] * np.sqrt(r[1])
                else:
                    assert np.abs(v[k] - r).max() <= tol * np.abs(r).max()
    for diag in "NU":
        for norm, ldv in (("M", 1), ("O", n), ("I", n), ("F", 2)):
            vals = dev_zeros(batch * ldv, np.float64)
            assert tr(ord(norm), ord(uplo), ord(diag), n, n, dA.p, n, vals.data_ptr(), ldv, batch, stream()) == 0
            sync()
            v = vals.cpu().numpy().reshape(batch, ldv)
            for k, a in enumerate(A):
                r = o.trnorm(norm, uplo, diag, a)
                if norm == "F":
                    assert abs(v[k, 0] * np.sqrt(v[k, 1]) - r[0] * np.sqrt(r[1])) <= tol * r[0] * np.sqrt(r[1])
                else:
                    assert np.abs(v[k] - r).max() <= tol * np.abs(np.atleast_1d(r)).max()
    # off-diagonal tile of a symmetric matrix: column sums then row sums
    m2 = n - 10
    B = rng_tiles(rng, batch, m2, n, "d"); dB = DevTiles(B)
    od = fn("sb200_synorm_offdiag_batched_d", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    vals = dev_zeros(batch * (m2 + n), np.float64)
    assert od(ord("O"), m2, n, dB.p, m2, vals.data_ptr(), m2 + n, batch, stream()) == 0
    sync()
    v = vals.cpu().numpy().reshape(batch, m2 + n)
    for k, b in enumerate(B):
        assert np.abs(v[k, :n] - np.abs(b).sum(axis=0)).max() <= tol * n
        assert np.abs(v[k, n:] - np.abs(b).sum(axis=1)).max() <= tol * n
