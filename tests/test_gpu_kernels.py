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


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
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
    opc = "C" if t in "cz" else "T"
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
        if t in "cz":
            d = np.arange(n); exp[d, d] = exp[d, d].real
        assert np.abs(x - exp).max() <= tol * np.abs(r).max()


def _trsm_cases():
    """every side / uplo / op / diag / layout for the four types; op 'C' only where it differs from 'T' (complex), the
    production tile size 512 x 512 for d (all ops) and z (N, C)"""
    import itertools
    out = []
    for t, layout, side, uplo, op, diag, (m, n) in itertools.product("dszc", "CR", "LR", "LU", "NTC", "NU",
                                                                     [(64, 48), (200, 136), (77, 53), (512, 512)]):
        if t in "ds" and op == "C":
            continue
        if (m, n) == (512, 512) and (t in "sc" or (op == "T" and t == "z")):
            continue
        out.append((t, layout, side, uplo, op, diag, m, n))
    return out


@pytest.mark.parametrize("t,layout,side,uplo,op,diag,m,n", _trsm_cases())
def test_trsm_batched(t, layout, side, uplo, op, diag, m, n):
    rng = np.random.default_rng(4)
    batch = 2
    na = m if side == "L" else n
    T = rng.random((na, na)) / na + np.eye(na) * (1 + rng.random(na))
    if t in "cz":
        T = T + 1j * rng.random((na, na)) / na
    T = T.astype(NP[t])
    B = rng_tiles(rng, batch, m, n, t)
    alpha = 0.7 - 0.2j if t in "cz" else 0.7
    wide = np.complex128 if t in "cz" else np.float64
    ref = [o.trsm_tile(side, uplo, op, diag, alpha, T.astype(wide), b.astype(wide)) for b in B]
    # row-major storage of X == column-major storage of X^T
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


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
@pytest.mark.parametrize("n", [37, 64, 200, 512])
def test_potrf_tile(t, n):
    rng = np.random.default_rng(5)
    G = rng.random((n, n))
    if t in "cz":
        G = G + 1j * rng.random((n, n))
    A = (G @ G.conj().T + n * np.eye(n)).astype(NP[t])
    dA = DevTiles([A.copy()])
    info = dev_zeros(1, np.int32)
    f = fn(f"sb200_potrf_tile_{t}", [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr])
    assert f(ord("L"), n, dA.t[0].data_ptr(), n, info.data_ptr(), None, stream()) == 0
    out = dA.get()[0]
    assert int(info.cpu()[0]) == 0
    ref = np.linalg.cholesky(A.astype(np.complex128 if t in "cz" else np.float64))
    assert np.abs(np.tril(out) - ref).max() <= 50 * EPS[t] * np.abs(ref).max()
    assert np.array_equal(np.triu(out, 1), np.triu(A, 1))         # strict upper triangle untouched


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
def test_syrk_and_single_tile_rank_k(t):
    """blas::batch::syrk / blas::syrk / blas::herk single-tile forms (device_batch_syrk.cc, device_syrk.cc,
    device_herk.cc): complex syrk keeps a complex diagonal, herk forces it real."""
    rng = np.random.default_rng(31)
    n, k, batch = 150, 40, 2
    A = rng_tiles(rng, batch, n, k, t); C = rng_tiles(rng, batch, n, n, t)
    alpha, beta = (0.5 + 0.25j, 1.5 - 1j) if t in "cz" else (0.5, 1.5)
    dA, dC = DevTiles(A), DevTiles(C)
    f = fn(f"sb200_syrk_batched_{t}", [c_int] * 3 + [c_i64] * 2 + [SC[t], c_ptr, c_i64, SC[t], c_ptr, c_i64, c_i64, c_ptr])
    assert f(ord("C"), ord("L"), ord("N"), n, k, scal(t, alpha), dA.p, n, scal(t, beta), dC.p, n, batch, stream()) == 0
    mask = np.tril(np.ones((n, n), bool))
    tol = 3 * np.sqrt(k) * EPS[t] * 4
    for x, a, c in zip(dC.get(), A, C):
        r = alpha * (a @ a.T) + beta * c
        assert np.abs(x - np.where(mask, r, c)).max() <= tol * np.abs(r).max()
    # single tile herk, upper, op = C (A is k x n)
    At = rng_tiles(rng, 1, k, n, t); C1 = rng_tiles(rng, 1, n, n, t)
    dA1, dC1 = DevTiles(At), DevTiles(C1)
    g = fn(f"sb200_herk_{t}", [c_int] * 3 + [c_i64] * 2 + [RSC[t], c_ptr, c_i64, RSC[t], c_ptr, c_i64, c_ptr])
    assert g(ord("C"), ord("U"), ord("C"), n, k, RSC[t](-1.0), dA1.t[0].data_ptr(), k, RSC[t](1.0), dC1.t[0].data_ptr(), n, stream()) == 0
    r = -1.0 * (At[0].conj().T @ At[0]) + C1[0]
    exp = np.where(mask.T, r, C1[0])
    if t in "cz":
        d = np.arange(n); exp[d, d] = exp[d, d].real
    assert np.abs(dC1.get()[0] - exp).max() <= tol * np.abs(r).max()
    # complex herk rejects op = T, complex syrk rejects op = C (blas::batch::herk_check semantics)
    if t in "cz":
        assert g(ord("C"), ord("U"), ord("T"), n, k, RSC[t](-1.0), dA1.t[0].data_ptr(), k, RSC[t](1.0), dC1.t[0].data_ptr(), n, stream()) == -1


@pytest.mark.parametrize("t", ["d", "z"])
def test_gemm_strided(t):
    """blas::gemm(queue) single GEMM and strided batches (device_gemm.cc): base + t * stride addressing."""
    import torch
    rng = np.random.default_rng(32)
    m, n, k, batch = 96, 80, 40, 3
    dt = NP[t]
    mk = lambda r, c: (rng.random((batch, c, r)) + (1j * rng.random((batch, c, r)) if t == "z" else 0)).astype(dt)
    A, B, C = mk(m, k), mk(k, n), mk(m, n)          # [t] holds the column-major tile as a (cols, rows) array
    tA, tB, tC = (torch.from_numpy(x.view(np.float64)).cuda() for x in (A, B, C))
    alpha, beta = (2.0 - 1j, 0.5 + 0.5j) if t == "z" else (2.0, 0.5)
    f = fn(f"sb200_gemm_strided_{t}", [c_int] * 3 + [c_i64] * 3 + [SC[t], c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, SC[t], c_ptr, c_i64, c_i64, c_i64, c_ptr])
    assert f(ord("C"), ord("N"), ord("N"), m, n, k, scal(t, alpha), tA.data_ptr(), m, m * k, tB.data_ptr(), k, k * n,
             scal(t, beta), tC.data_ptr(), m, m * n, batch, stream()) == 0
    sync()
    out = tC.cpu().numpy().view(dt).reshape(batch, n, m)
    for i in range(batch):
        r = alpha * (A[i].T @ B[i].T) + beta * C[i].T
        assert np.abs(out[i].T - r).max() <= 3 * np.sqrt(k) * EPS[t] * 4 * np.abs(r).max()


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
@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
@pytest.mark.parametrize("m,n", [(64, 48), (257, 129), (512, 512), (1, 7)])
def test_geadd_gescale_geset(t, m, n):
    rng = np.random.default_rng(7)
    batch = 3
    A = rng_tiles(rng, batch, m, n, t); B = rng_tiles(rng, batch, m, n, t)
    alpha, beta = (1.5 - 0.5j, 0.25 + 2j) if t in "cz" else (1.5, 0.25)
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


@pytest.mark.parametrize("pair", ["dd", "ds", "sd", "ss", "zz", "zc", "cz", "cc", "sc", "dz"])
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


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
@pytest.mark.parametrize("m,n", [(64, 64), (100, 36), (33, 257), (512, 512)])
def test_transposes(t, m, n):
    rng = np.random.default_rng(11)
    batch = 2
    A = rng_tiles(rng, batch, m, n, t)
    dA = DevTiles(A); dT = DevTiles([np.zeros((n, m), NP[t], order="F") for _ in range(batch)])
    conj = 1 if t in "cz" else 0
    f = fn(f"sb200_transpose_batched_{t}", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    assert f(conj, m, n, dA.p, m, dT.p, n, batch, stream()) == 0
    for x, a in zip(dT.get(), A):
        assert np.array_equal(x, a.conj().T if conj else a.T)
    # single-tile out-of-place entry (device::transpose, device_transpose.cu): tile 0 -> tile 1 of dT
    g = fn(f"sb200_transpose_{t}", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr])
    assert g(0, m, n, dA.t[1].data_ptr(), m, dT.t[0].data_ptr(), n, stream()) == 0
    assert np.array_equal(dT.get()[0], A[1].T)
    if m == n:
        f = fn(f"sb200_transpose_inplace_batched_{t}", [c_int, c_i64, c_ptr, c_i64, c_i64, c_ptr])
        assert f(0, n, dA.p, n, batch, stream()) == 0
        for x, a in zip(dA.get(), A):
            assert np.array_equal(x, a.T)
        g = fn(f"sb200_transpose_inplace_{t}", [c_int, c_i64, c_ptr, c_i64, c_ptr])
        assert g(conj, n, dA.t[0].data_ptr(), n, stream()) == 0
        assert np.array_equal(dA.get()[0], A[0].conj() if conj else A[0])


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
def test_single_tile_entries(t):
    """The reference's un-batched device::geadd / gescale / geset / tzset (device_geadd.cu:121-146,
    device_gescale.cu:79-110, device_geset.cu:98-125, device_tzset.cu:98-130)."""
    rng = np.random.default_rng(21)
    m, n = 130, 70
    A = rng_tiles(rng, 1, m, n, t); B = rng_tiles(rng, 1, m, n, t)
    dA, dB = DevTiles(A), DevTiles(B)
    alpha, beta = (1.5 - 0.5j, 0.25 + 2j) if t in "cz" else (1.5, 0.25)
    f = fn(f"sb200_geadd_{t}", [c_i64, c_i64, SC[t], c_ptr, c_i64, SC[t], c_ptr, c_i64, c_ptr])
    assert f(m, n, scal(t, alpha), dA.t[0].data_ptr(), m, scal(t, beta), dB.t[0].data_ptr(), m, stream()) == 0
    r = o.geadd(alpha, A[0], beta, B[0])
    assert np.abs(dB.get()[0] - r).max() <= 4 * EPS[t] * np.abs(r).max()
    f = fn(f"sb200_gescale_{t}", [c_i64, c_i64, SC[t], SC[t], c_ptr, c_i64, c_ptr])
    assert f(m, n, scal(t, 3.0), scal(t, 2.0), dA.t[0].data_ptr(), m, stream()) == 0
    assert np.abs(dA.get()[0] - o.gescale(3.0, 2.0, A[0])).max() <= 6 * EPS[t] * np.abs(A[0]).max()
    f = fn(f"sb200_geset_{t}", [c_int, c_i64, c_i64, SC[t], SC[t], c_ptr, c_i64, c_ptr])
    assert f(ord("G"), m, n, scal(t, 0.5), scal(t, 9.0), dA.t[0].data_ptr(), m, stream()) == 0
    assert np.array_equal(dA.get()[0], o.geset(0.5, 9.0, m, n, NP[t]))
    assert f(ord("L"), m, n, scal(t, 2.0), scal(t, 1.0), dA.t[0].data_ptr(), m, stream()) == 0
    assert np.array_equal(dA.get()[0], o.tzset("L", 2.0, 1.0, o.geset(0.5, 9.0, m, n, NP[t])))


@pytest.mark.parametrize("t", ["z", "c"])
def test_gescale_row_col_real_scales_on_complex_tiles(t):
    rng = np.random.default_rng(22)
    m, n, batch = 96, 80, 2
    A = rng_tiles(rng, batch, m, n, t)
    rt = REAL[t]
    R = [rng.random((m, 1)).astype(rt) for _ in range(batch)]; Cc = [rng.random((n, 1)).astype(rt) for _ in range(batch)]
    f = fn(f"sb200_gescale_row_col_real_batched_{t}", [c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr])
    dA, dR, dC = DevTiles(A), DevTiles(R), DevTiles(Cc)
    assert f(ord("B"), m, n, dR.p, dC.p, dA.p, m, batch, stream()) == 0
    for x, a, r, c in zip(dA.get(), A, R, Cc):
        ref = o.gescale_row_col("B", r[:, 0], c[:, 0], a)
        assert np.abs(x - ref).max() <= 4 * EPS[t] * np.abs(ref).max()


@pytest.mark.parametrize("t", ["d", "s", "z", "c"])
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
            r = o.genorm(norm, a.astype(np.complex128 if t in "cz" else np.float64))
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
