"""oracle/slate_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU (numpy) restatement of the reference algorithms on the hot path.  It is the
checker for the CUDA path; it is never imported by the product package
(slate_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it.

PINNED: every function below is checked against the UNMODIFIED reference built by
oracle/build_ref.sh (oracle/_ref/ref_dump, the reference's own HostTask path; on process
grids oracle/build_ref_mp.sh -> oracle/_ref/ref_dump_mp under oracle/mprun.py) --
bit-exact for the Philox generator and the pivot vectors, to a few ulp for the
floating-point factors -- by tests/test_oracle.py, using the committed fixtures in
tests/golden/ (made by tests/golden/make_golden.py) and, when oracle/_ref is
present, live runs of ref_dump.

Each function cites the reference file:line it restates.  The arithmetic that the
reference delegates to host BLAS/LAPACK (dgemm, dsyrk, dtrsm, dpotrf; OpenBLAS in
this image) is done here with numpy/scipy calls of the same mathematical
operation, tile by tile and in the reference's order of tile operations, so
results agree to rounding, not bitwise.
"""
from __future__ import annotations

import functools

import numpy as np

# ----------------------------------------------------------------------------
# matgen: Philox-2x64 keyed on (global i, global j, seed)
# reference: matgen/random.cc:54-77 (philox_2x64), :83-91 (rand_to_real), :96-159
# ----------------------------------------------------------------------------
_PHILOX_MULT = np.uint64(0x9E3779B97F4A7C15)
_PHILOX_SEED_INC = np.uint64(0xD2B74407B1CE6E93)
_M32 = np.uint64(0xFFFFFFFF)


def _mulhilo64(a: np.ndarray, b: np.uint64):
    """128-bit product of uint64 array `a` with scalar `b` -> (lo, hi), wrap-around uint64."""
    a = a.astype(np.uint64, copy=False)
    a_lo, a_hi = a & _M32, a >> np.uint64(32)
    b_lo, b_hi = b & _M32, b >> np.uint64(32)
    ll = a_lo * b_lo
    lh = a_lo * b_hi
    hl = a_hi * b_lo
    hh = a_hi * b_hi
    mid = (ll >> np.uint64(32)) + (lh & _M32) + (hl & _M32)
    lo = (ll & _M32) | ((mid & _M32) << np.uint64(32))
    hi = hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))
    return lo, hi


def philox_2x64(i: np.ndarray, j: np.ndarray, seed: int):
    """10 rounds of Philox-2x64 on counter (i, j), key `seed` (matgen/random.cc:54-77)."""
    with np.errstate(over="ignore"):
        s0 = np.asarray(i, dtype=np.int64).astype(np.uint64)
        s1 = np.asarray(j, dtype=np.int64).astype(np.uint64)
        s0, s1 = np.broadcast_arrays(s0, s1)
        key = np.uint64(np.int64(seed).astype(np.uint64))
        for rnd in range(10):
            if rnd != 0:
                key = key + _PHILOX_SEED_INC
            lo, hi = _mulhilo64(s1, _PHILOX_MULT)
            s0, s1 = lo, hi ^ key ^ s0
    return s0, s1


def _to_real(bits: np.ndarray, dtype):
    """uniform [0,1) from the top `digits` bits (matgen/random.cc:83-91)."""
    digits = np.finfo(dtype).nmant + 1
    return (bits >> np.uint64(64 - digits)).astype(dtype) / dtype(2.0 ** digits)


def generate(kind: str, m: int, n: int, seed: int, dtype=np.float64,
             i0: int = 0, j0: int = 0, n_global: int | None = None) -> np.ndarray:
    """Dense m-by-n block starting at global (i0, j0) of the reference `rand` /
    `rand_dominant` test matrix (matgen/generate_type_rand.hh:28-79: uniform [0,1),
    dominant adds n to the diagonal; complex: (re, im) from the two Philox words).
    The parametrised GPU tests ask for the same few matrices over and over (2.9 s for 2048 x 2048 in numpy): the last
    few results are kept and a fresh copy is handed out."""
    return _generate_cached(kind, int(m), int(n), int(seed), np.dtype(dtype).str, int(i0), int(j0),
                            None if n_global is None else int(n_global)).copy(order="F")


@functools.lru_cache(maxsize=12)
def _generate_cached(kind, m, n, seed, dtype, i0, j0, n_global):
    dtype = np.dtype(dtype)
    real = np.float32 if dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    ii = np.arange(i0, i0 + m, dtype=np.int64)[:, None]
    jj = np.arange(j0, j0 + n, dtype=np.int64)[None, :]
    b0, b1 = philox_2x64(ii, jj, seed)
    A = _to_real(b0, real)
    if dtype.kind == "c":
        A = A + 1j * _to_real(b1, real)
    A = np.asfortranarray(A.astype(dtype))
    if kind == "rand_dominant":
        ng = n if n_global is None else n_global
        d = np.arange(max(i0, j0), min(i0 + m, j0 + n))
        A[d - i0, d - j0] += ng
    elif kind != "rand":
        raise ValueError(kind)
    return A


# ----------------------------------------------------------------------------
# tile helpers
# ----------------------------------------------------------------------------
def _tiles(n: int, nb: int):
    return [(s, min(s + nb, n)) for s in range(0, n, nb)]


def cabs1(x):
    """|re| + |im| (blaspp cabs1; used by the pivot search, Tile_getrf.hh:213)."""
    x = np.asarray(x)
    return np.abs(x.real) + np.abs(x.imag) if np.iscomplexobj(x) else np.abs(x)


# ----------------------------------------------------------------------------
# gemm: C = alpha A B + beta C as gemmC does it (src/gemmC.cc:124-182): beta applied
# with the k = 0 block, then one rank-nb update per block column k of A.
# tile op: include/slate/Tile_blas.hh:29-98 -> blas::gemm
# ----------------------------------------------------------------------------
def gemm(alpha, A, B, beta, C, nb: int, opa: str = "N", opb: str = "N"):
    """C = alpha op(A) op(B) + beta C, one block column of op(A) per step (src/gemmC.cc:124-182); op = N / T / C are the
    (conjugate-)transposed views slate::gemm is handed (A stored k x m, B stored n x k)."""
    view = {"N": lambda x: x, "T": lambda x: x.T, "C": lambda x: x.conj().T}
    A, B = view[opa](np.asarray(A)), view[opb](np.asarray(B))
    C = np.array(C, order="F", copy=True)
    kt = _tiles(A.shape[1], nb)
    for idx, (k0, k1) in enumerate(kt):
        b = beta if idx == 0 else 1.0
        for (j0, j1) in _tiles(C.shape[1], nb):
            for (i0, i1) in _tiles(C.shape[0], nb):
                C[i0:i1, j0:j1] = alpha * (A[i0:i1, k0:k1] @ B[k0:k1, j0:j1]) + b * C[i0:i1, j0:j1]
    return C


# ----------------------------------------------------------------------------
# herk: lower C = alpha A A^H + beta C (src/herk.cc:25-162; internal_herk.cc:47-110):
# diagonal tiles by tile herk (only the stored triangle is defined), off-diagonal by gemm.
# ----------------------------------------------------------------------------
def herk(alpha, A, beta, C, nb: int, lower: bool = True):
    C = np.array(C, order="F", copy=True)
    n = C.shape[0]
    kt = _tiles(A.shape[1], nb)
    for idx, (k0, k1) in enumerate(kt):
        b = beta if idx == 0 else 1.0
        for (j0, j1) in _tiles(n, nb):
            for (i0, i1) in _tiles(n, nb):
                if (lower and i0 < j0) or (not lower and i0 > j0):
                    continue
                upd = alpha * (A[i0:i1, k0:k1] @ A[j0:j1, k0:k1].conj().T) + b * C[i0:i1, j0:j1]
                if i0 == j0:
                    mask = np.tril(np.ones_like(upd, dtype=bool)) if lower else np.triu(np.ones_like(upd, dtype=bool))
                    blk = C[i0:i1, j0:j1]
                    blk[mask] = upd[mask]
                    if np.iscomplexobj(blk):
                        di = np.arange(blk.shape[0])
                        blk[di, di] = blk[di, di].real
                else:
                    C[i0:i1, j0:j1] = upd
    return C


# ----------------------------------------------------------------------------
# potrf: right-looking tile Cholesky, lower (src/potrf.cc:84-195):
#   per k: potrf(A_kk) (internal_potrf.cc -> lapack::potrf), trsm of the column
#   (internal_trsm.cc: Right, Lower, ConjTrans, NonUnit), herk/gemm trailing update.
# Returns (L with the strict upper triangle zeroed, info).
# ----------------------------------------------------------------------------
def potrf(A, nb: int):
    from scipy.linalg import solve_triangular
    A = np.array(A, order="F", copy=True)
    n = A.shape[0]
    tl = _tiles(n, nb)
    info = 0
    for k, (k0, k1) in enumerate(tl):
        Akk = np.tril(A[k0:k1, k0:k1])
        Akk = Akk + np.tril(Akk, -1).conj().T
        try:
            L = np.linalg.cholesky(Akk)
        except np.linalg.LinAlgError:
            # first non-positive leading minor, 1-based global index (LAPACK info)
            for r in range(1, k1 - k0 + 1):
                try:
                    np.linalg.cholesky(Akk[:r, :r])
                except np.linalg.LinAlgError:
                    info = k0 + r
                    break
            break
        A[k0:k1, k0:k1] = L
        if k1 < n:
            # A(i,k) <- A(i,k) L^{-H}
            A[k1:, k0:k1] = solve_triangular(L, A[k1:, k0:k1].conj().T, lower=True).conj().T
            P = A[k1:, k0:k1]
            for (j0, j1) in tl[k + 1:]:
                for (i0, i1) in tl[k + 1:]:
                    if i0 < j0:
                        continue
                    A[i0:i1, j0:j1] -= P[i0 - k1:i1 - k1] @ P[j0 - k1:j1 - k1].conj().T
    return np.tril(A), info


# ----------------------------------------------------------------------------
# getrf: tile LU with partial pivoting.
#   panel  : src/internal/Tile_getrf.hh:160-447 (ib-blocked, pivot = first strict max of
#            cabs1 starting from the diagonal entry -- ties keep the LOWEST row; scale by
#            the reciprocal unless |pivot| < safe_min; zero pivot -> info = j+1, continue)
#   driver : src/getrf.cc:84-236 (permuteRows right, trsm Left/Lower/Unit, gemm update,
#            then permuteRows to the left)
# Returns (LU, pivots, info); pivots[k] is a list of (tileIndex, elementOffset) relative to
# the panel sub-matrix A(k:mt-1, k)  (include/slate/types.hh:84-105).
# ----------------------------------------------------------------------------
def getrf_panel(P: np.ndarray, diag_len: int, ib: int, nb_rows: int):
    """Factor the m-by-nb panel P in place; returns (pivot rows (panel-relative), info)."""
    m, nbc = P.shape
    safe_min = np.finfo(P.real.dtype).tiny
    piv = np.zeros(diag_len, dtype=np.int64)
    info = 0
    for k in range(0, diag_len, ib):
        kb = min(diag_len - k, ib)
        for j in range(k, k + kb):
            col = cabs1(P[j:, j])
            r = j + int(np.argmax(col))          # first maximum = lowest row among ties
            piv[j] = r
            if r != j:
                P[[j, r], :] = P[[r, j], :]
            pv = P[j, j]
            if cabs1(pv) >= safe_min:
                P[j + 1:, j] *= (1.0 / pv)
            elif pv != 0:
                P[j + 1:, j] /= pv
            elif info == 0:
                info = j + 1
            if k + kb > j + 1:
                P[j + 1:, j + 1:k + kb] -= np.outer(P[j + 1:, j], P[j, j + 1:k + kb])
        if k + kb < nbc:
            # trsm on the top rows of the stripe, then rank-kb update to the right
            L = np.tril(P[k:k + kb, k:k + kb], -1) + np.eye(kb, dtype=P.dtype)
            P[k:k + kb, k + kb:] = np.linalg.solve(L, P[k:k + kb, k + kb:]) if kb > 1 else P[k:k + kb, k + kb:]
            P[k + kb:, k + kb:] -= P[k + kb:, k:k + kb] @ P[k:k + kb, k + kb:]
    return piv, info


_GETRF_MEMO: dict = {}


def getrf(A, nb: int, ib: int = 16):
    """Blocked LU with the reference's pivot rule.  Memoised on the input's bytes (the parametrised GPU tests factor the
    same seeded matrix once per kernel variant); callers get copies."""
    import hashlib
    Af = np.asfortranarray(A)
    key = (hashlib.blake2b(Af.tobytes(order="F"), digest_size=16).digest(), Af.shape, Af.dtype.str, int(nb), int(ib))
    hit = _GETRF_MEMO.get(key)
    if hit is None:
        if len(_GETRF_MEMO) >= 12:
            _GETRF_MEMO.pop(next(iter(_GETRF_MEMO)))
        hit = _GETRF_MEMO[key] = _getrf_impl(Af, nb, ib)
    LU, piv, info = hit
    return LU.copy(order="F"), [list(c) for c in piv], info


def _getrf_impl(A, nb: int, ib: int = 16):
    from scipy.linalg import solve_triangular
    A = np.array(A, order="F", copy=True)
    m, n = A.shape
    info = 0
    pivots = []
    kk = 0
    for k in range(min(-(-m // nb), -(-n // nb))):
        r0, c0 = k * nb, k * nb
        r1, c1 = min(r0 + nb, m), min(c0 + nb, n)
        diag_len = min(r1 - r0, c1 - c0)
        panel = A[r0:, c0:c1]
        piv, iinfo = getrf_panel(panel, diag_len, ib, nb)
        if info == 0 and iinfo > 0:
            info = kk + iinfo
        pivots.append([(int(p // nb), int(p % nb)) for p in piv])
        # apply the row interchanges to the columns right and left of the panel
        for j, p in enumerate(piv):
            if p != j:
                A[[r0 + j, r0 + p], c1:] = A[[r0 + p, r0 + j], c1:]
                A[[r0 + j, r0 + p], :c0] = A[[r0 + p, r0 + j], :c0]
        if c1 < n:
            Lkk = A[r0:r1, c0:c1][:, :diag_len]
            A[r0:r1, c1:] = solve_triangular(Lkk[:diag_len], A[r0:r1, c1:], lower=True, unit_diagonal=True)
            if r1 < m:
                for (j0, j1) in _tiles(n, nb)[k + 1:]:
                    A[r1:, j0:j1] -= A[r1:, c0:c1] @ A[r0:r1, j0:j1]
        kk += c1 - c0
    return A, pivots, info


def pivots_to_perm(pivots, m: int, nb: int) -> np.ndarray:
    """Row permutation p such that (P A)[i] = A[p[i]] for slate::Pivots semantics."""
    perm = np.arange(m)
    for k, col in enumerate(pivots):
        r0 = k * nb
        for j, (ti, off) in enumerate(col):
            a, b = r0 + j, r0 + ti * nb + off
            if a != b:
                perm[[a, b]] = perm[[b, a]]
    return perm


# ----------------------------------------------------------------------------
# trsm with one triangular tile against a block row/column (internal_trsm.cc:40-92)
# ----------------------------------------------------------------------------
def trsm_tile(side: str, uplo: str, op: str, diag: str, alpha, T, B):
    from scipy.linalg import solve_triangular
    lower = uplo == "L"
    Tm = np.tril(T) if lower else np.triu(T)
    trans = {"N": 0, "T": 1, "C": 2}[op]
    unit = diag == "U"
    if side == "L":
        return solve_triangular(Tm, alpha * B, lower=lower, trans=trans, unit_diagonal=unit)
    # X op(T) = alpha B  <=>  op(T)^T X^T = alpha B^T
    if op == "N":
        Xt = solve_triangular(Tm, (alpha * B).T, lower=lower, trans=1, unit_diagonal=unit)
    elif op == "T":
        Xt = solve_triangular(Tm, (alpha * B).T, lower=lower, trans=0, unit_diagonal=unit)
    else:
        Xt = solve_triangular(Tm.conj(), (alpha * B).T, lower=lower, trans=0, unit_diagonal=unit)
    return Xt.T


# ----------------------------------------------------------------------------
# solve path: block-row sweeps of work::trsm (src/work/work_trsm.cc:60-230), potrs
# (src/potrs.cc:54-77), getrs (src/getrs.cc:25-66), hemm (src/hemmC.cc, Left/Lower) and the
# Hermitian inf-norm (src/norm.cc -> internal_henorm.cc: row sums over the implied full matrix)
# ----------------------------------------------------------------------------
def tri_sweep(Tm, B, nb: int, lower: bool, op: str = "N", unit: bool = False):
    """B <- op(T)^{-1} B, one diagonal-tile solve + gemm update of the remaining block rows per step."""
    from scipy.linalg import solve_triangular
    B = np.array(B, order="F", copy=True)
    n = Tm.shape[0]
    tl = _tiles(n, nb)
    trans = op != "N"
    M = Tm if not trans else (Tm.conj().T if op == "C" else Tm.T)      # op(T) as a math matrix
    eff_lower = lower != trans
    order = tl if eff_lower else tl[::-1]
    for (k0, k1) in order:
        B[k0:k1] = solve_triangular(M[k0:k1, k0:k1], B[k0:k1], lower=eff_lower, unit_diagonal=unit)
        rows = [(i0, i1) for (i0, i1) in tl if (i0 > k0 if eff_lower else i0 < k0)]
        for (i0, i1) in rows:
            B[i0:i1] -= M[i0:i1, k0:k1] @ B[k0:k1]
    return B


def tri_sweep_right(Tm, B, nb: int, lower: bool, op: str = "N", unit: bool = False):
    """B <- B op(T)^{-1}: the block-column sweep slate::trsm runs for Side::Right (src/trsm.cc -> work::trsm on the
    (conjugate-)transposed views, src/work/work_trsm.cc:78-98): one diagonal-tile solve of block column k + gemm update
    of the block columns that still wait, backward when op(T) is lower, forward when it is upper."""
    from scipy.linalg import solve_triangular
    B = np.array(B, order="F", copy=True)
    tl = _tiles(Tm.shape[0], nb)
    trans = op != "N"
    M = Tm if not trans else (Tm.conj().T if op == "C" else Tm.T)      # op(T) as a math matrix
    eff_lower = lower != trans
    order = tl[::-1] if eff_lower else tl
    for (k0, k1) in order:
        # X_k M_kk = B_k  <=>  M_kk^T X_k^T = B_k^T
        B[:, k0:k1] = solve_triangular(M[k0:k1, k0:k1], B[:, k0:k1].T, lower=eff_lower, trans=1, unit_diagonal=unit).T
        cols = [(j0, j1) for (j0, j1) in tl if (j0 < k0 if eff_lower else j0 > k0)]
        for (j0, j1) in cols:
            B[:, j0:j1] -= B[:, k0:k1] @ M[k0:k1, j0:j1]
    return B


def trsm(alpha, T, B, nb: int, side: str = "L", lower: bool = True, op: str = "N", unit: bool = False):
    """slate::trsm(side, alpha, op(T), B): B <- alpha op(T)^{-1} B (Left) or alpha B op(T)^{-1} (Right), T the lower / upper
    triangle of the tile matrix (src/trsm.cc; alpha is applied to B before the sweep, src/work/work_trsm.cc:100-130)."""
    Tm = (np.tril(T) if lower else np.triu(T)).astype(np.result_type(T, B), copy=True)
    Bs = alpha * np.asarray(B)
    return (tri_sweep if side == "L" else tri_sweep_right)(Tm, Bs, nb, lower, op, unit)


def potrs(L, B, nb: int):
    """Solve A X = B with A = L L^H (lower factor from potrf)."""
    Y = tri_sweep(np.tril(L), B, nb, lower=True, op="N")
    return tri_sweep(np.tril(L), Y, nb, lower=True, op="C")


def getrs(LU, pivots, B, nb: int, op: str = "N"):
    """Solve op(A) X = B with P A = L U from getrf (pivots as slate::Pivots; src/getrs.cc:81-112): NoTrans = permute,
    L sweep, U sweep; transposed = op(U) sweep, op(L) sweep, inverse permutation."""
    perm = pivots_to_perm(pivots, LU.shape[0], nb)
    Lm = np.tril(LU, -1) + np.eye(LU.shape[0], dtype=LU.dtype)
    if op == "N":
        Y = tri_sweep(Lm, np.asarray(B)[perm], nb, lower=True, unit=True)
        return tri_sweep(np.triu(LU), Y, nb, lower=False)
    Y = tri_sweep(np.triu(LU), B, nb, lower=False, op=op)
    Xh = tri_sweep(Lm, Y, nb, lower=True, op=op, unit=True)
    X = np.empty_like(Xh)
    X[perm] = Xh                                   # X = P^T Xhat
    return X


# ----------------------------------------------------------------------------
# getrf_tntpiv: LU with tournament pivoting (CALU).  src/getrf_tntpiv.cc:22-395 (driver),
# src/internal/internal_getrf_tntpiv.cc:357-640 (panel: local LU per rank, binary tree of pairwise LUs of the
# stacked candidate rows, permutation_to_sequential_pivot :42-120), src/internal/Tile_getrf_tntpiv.hh:69-300
# (the LU used at every node: same pivot rule as tile::getrf).
#
#   per panel k:  ranks = the process rows that own tiles of block column k (tileRank(i, k) = i % p + ...), ordered by
#     their first tile.  Stage 0: every rank factors a COPY of its own rows with partial pivoting and keeps the ORIGINAL
#     rows that ended in its first piv_len positions.  Tree: at level l the rank at index x (x % 2^(l+1) == 0) stacks
#     its candidate tile on the one of index x + 2^l, factors a copy, and keeps the original rows of the winners.
#     The last LU (rank index 0) gives the factored diagonal tile and the nb winning rows; their sequence of row
#     interchanges follows from where each winner sits when its turn comes.
#   driver: interchanges applied to the whole block row range, A(k,k) <- factored tile, A(i,k) <- A(i,k) U_kk^-1 below
#     it (trsm Right/Upper/NonUnit, :171-184), then the row solve and trailing update of getrf.
#
# PINNED: for ranks = 1 by tests/golden/getrf_tntpiv_d*.npz (the one-rank reference, oracle/_ref/ref_dump); for ranks > 1 by
# tests/golden/grid_getrf_tntpiv_d_*.npz, written by the unmodified reference RUNNING ON 2x1, 3x1, 4x1 and 2x4 PROCESS GRIDS
# (oracle/_ref/ref_dump_mp: the reference built against the multi-process MPI replacement oracle/mpi_mp, oracle/mprun.py) --
# identical pivots, factors to 1e-13 (tests/test_oracle.py).
# With one rank the pivots are those of partial pivoting; the factor differs from getrf's in rounding only.
# ----------------------------------------------------------------------------
def tnt_winners_to_sequential(winners, m_p: int):
    """Sequential interchanges (position j <-> piv[j]) that bring original row winners[j] to position j, and the final
    arrangement row_at (row_at[x] = original row at position x).  internal_getrf_tntpiv.cc:42-120."""
    row_at = np.arange(m_p)
    pos_of = np.arange(m_p)
    piv = np.zeros(len(winners), dtype=np.int64)
    for j, w in enumerate(winners):
        x = int(pos_of[w])
        assert x >= j
        piv[j] = x
        rj = int(row_at[j])
        row_at[j], row_at[x] = w, rj
        pos_of[w], pos_of[rj] = j, x
    return piv, row_at


def tnt_panel(P: np.ndarray, nb: int, k: int, ranks: int, ib: int):
    """Tournament over the m_p x kw panel P (ORIGINAL rows; tile t of the panel is global tile row k + t).
    Returns (sequential pivots (panel-relative rows), factored top tile (diag rows x kw), info)."""
    m_p, kw = P.shape
    ntile = -(-m_p // nb)
    tile_rows = [np.arange(t * nb, min((t + 1) * nb, m_p)) for t in range(ntile)]
    order = []                                   # ranks by first tile (rank_rows, :456-470)
    for t in range(ntile):
        r = (k + t) % ranks
        if r not in order:
            order.append(r)
    nranks = len(order)
    if nranks > 1:
        assert kw == nb, "tournament over several ranks restated for full-width panels only"
    cand = []                                    # per rank index: original panel rows of its candidate tile, in order
    info = 0
    top = None
    for x, r in enumerate(order):
        rows = np.concatenate([tile_rows[t] for t in range(ntile) if (k + t) % ranks == r])
        first_mb = len(tile_rows[[t for t in range(ntile) if (k + t) % ranks == r][0]])
        piv_len = min(first_mb, kw)
        W = np.array(P[rows], order="F", copy=True)
        piv, iinfo = getrf_panel(W, piv_len, ib, nb)
        ids = rows.copy()
        for j, pj in enumerate(piv):
            if pj != j:
                ids[[j, pj]] = ids[[pj, j]]
        cand.append(ids[:first_mb].copy())
        if x == 0:
            info = iinfo
            top = W[:first_mb].copy()
            seq = piv
    if nranks == 1:
        return seq, top, info
    nlevels = int(np.ceil(np.log2(nranks)))
    step = 1
    for level in range(nlevels):
        for x in range(0, nranks, 2 * step):
            if x + step < nranks:
                ids = np.concatenate([cand[x], cand[x + step]])
                mb1 = len(cand[x])
                W = np.array(P[ids], order="F", copy=True)
                piv, iinfo = getrf_panel(W, min(mb1, kw), ib, nb)
                for j, pj in enumerate(piv):
                    if pj != j:
                        ids[[j, pj]] = ids[[pj, j]]
                cand[x] = ids[:mb1].copy()
                if x == 0:
                    if info == 0:
                        info = iinfo
                    top = W[:mb1].copy()
        step *= 2
    diag_len = min(len(tile_rows[0]), kw)
    seq, _ = tnt_winners_to_sequential(cand[0][:diag_len], m_p)
    return seq, top, info


def getrf_tntpiv(A, nb: int, ib: int = 16, ranks: int = 1):
    from scipy.linalg import solve_triangular
    A = np.array(A, order="F", copy=True)
    m, n = A.shape
    info = 0
    pivots = []
    kk = 0
    for k in range(min(-(-m // nb), -(-n // nb))):
        r0, c0 = k * nb, k * nb
        r1, c1 = min(r0 + nb, m), min(c0 + nb, n)
        diag_len = min(r1 - r0, c1 - c0)
        piv, top, iinfo = tnt_panel(A[r0:, c0:c1], nb, k, ranks, ib)
        if info == 0 and iinfo > 0:
            info = kk + iinfo
        pivots.append([(int(p // nb), int(p % nb)) for p in piv[:diag_len]])
        for j, p in enumerate(piv[:diag_len]):          # permuteRows on every block column, the panel's included
            if p != j:
                A[[r0 + j, r0 + p], :] = A[[r0 + p, r0 + j], :]
        A[r0:r1, c0:c1] = top
        if r1 < m:                                      # A(i, k) <- A(i, k) U_kk^-1
            U = np.triu(top[:diag_len, :diag_len])
            A[r1:, c0:c0 + diag_len] = solve_triangular(U, A[r1:, c0:c0 + diag_len].T, lower=False, trans=1).T
        if c1 < n:
            Lkk = A[r0:r1, c0:c1][:, :diag_len]
            A[r0:r1, c1:] = solve_triangular(Lkk[:diag_len], A[r0:r1, c1:], lower=True, unit_diagonal=True)
            if r1 < m:
                for (j0, j1) in _tiles(n, nb)[k + 1:]:
                    A[r1:, j0:j1] -= A[r1:, c0:c1] @ A[r0:r1, j0:j1]
        kk += c1 - c0
    return A, pivots, info


# ----------------------------------------------------------------------------
# getrf_nopiv: A = L U without pivoting (src/getrf_nopiv.cc:25-190; internal_getrf_nopiv.cc + Tile_getrf_nopiv.hh):
# per step the diagonal tile is factored (ib-blocked, no search), the column below is solved against U_kk, the row to
# the right against L_kk, the rest updated.  The factors of LU without pivoting are unique, so the blocking only
# changes rounding; this restatement is the tile-level right-looking form.  info = first zero pivot + 1.
# ----------------------------------------------------------------------------
def getrf_nopiv(A, nb: int):
    A = np.array(A, order="F", copy=True)
    m, n = A.shape
    info = 0
    for (k0, k1) in _tiles(min(m, n), nb):
        for j in range(k0, k1):                                   # diagonal tile (and, equivalently, the column below it)
            piv = A[j, j]
            if piv == 0:
                info = info or j + 1
                continue
            A[j + 1:, j] /= piv
            A[j + 1:, j + 1:k1] -= np.outer(A[j + 1:, j], A[j, j + 1:k1])
        if k1 < n:
            L = np.tril(A[k0:k1, k0:k1], -1) + np.eye(k1 - k0)
            A[k0:k1, k1:] = np.linalg.solve(L, A[k0:k1, k1:])
            A[k1:, k1:] -= A[k1:, k0:k1] @ A[k0:k1, k1:]
    return A, info


# ----------------------------------------------------------------------------
# her2k: lower C = alpha A B^H + conj(alpha) B A^H + beta C (src/her2k.cc:27-170; internal_her2k.cc):
# one block column k of A and B per step, diagonal tiles by tile her2k (stored triangle, real diagonal),
# off-diagonal tiles by two gemms.  Real types: syr2k.
# ----------------------------------------------------------------------------
def her2k(alpha, A, B, beta, C, nb: int):
    C = np.array(C, order="F", copy=True)
    n = C.shape[0]
    for idx, (k0, k1) in enumerate(_tiles(A.shape[1], nb)):
        b = beta if idx == 0 else 1.0
        for (j0, j1) in _tiles(n, nb):
            for (i0, i1) in _tiles(n, nb):
                if i0 < j0:
                    continue
                upd = (alpha * (A[i0:i1, k0:k1] @ B[j0:j1, k0:k1].conj().T)
                       + np.conj(alpha) * (B[i0:i1, k0:k1] @ A[j0:j1, k0:k1].conj().T) + b * C[i0:i1, j0:j1])
                if i0 == j0:
                    mask = np.tril(np.ones_like(upd, dtype=bool))
                    blk = C[i0:i1, j0:j1]
                    blk[mask] = upd[mask]
                    if np.iscomplexobj(blk):
                        di = np.arange(blk.shape[0])
                        blk[di, di] = blk[di, di].real
                else:
                    C[i0:i1, j0:j1] = upd
    return C


# ----------------------------------------------------------------------------
# syrk / syr2k for complex-symmetric C (no conjugation; src/syrk.cc, src/syr2k.cc): alpha, beta are full scalars and
# the diagonal stays complex.  (For real types they coincide with herk / her2k.)
# ----------------------------------------------------------------------------
def syrk(alpha, A, beta, C, nb: int):
    C = np.array(C, order="F", copy=True)
    n = C.shape[0]
    for idx, (k0, k1) in enumerate(_tiles(A.shape[1], nb)):
        b = beta if idx == 0 else 1.0
        for (j0, j1) in _tiles(n, nb):
            for (i0, i1) in _tiles(n, nb):
                if i0 < j0:
                    continue
                upd = alpha * (A[i0:i1, k0:k1] @ A[j0:j1, k0:k1].T) + b * C[i0:i1, j0:j1]
                if i0 == j0:
                    mask = np.tril(np.ones_like(upd, dtype=bool))
                    C[i0:i1, j0:j1][mask] = upd[mask]
                else:
                    C[i0:i1, j0:j1] = upd
    return C


def syr2k(alpha, A, B, beta, C, nb: int):
    C = np.array(C, order="F", copy=True)
    n = C.shape[0]
    for idx, (k0, k1) in enumerate(_tiles(A.shape[1], nb)):
        b = beta if idx == 0 else 1.0
        for (j0, j1) in _tiles(n, nb):
            for (i0, i1) in _tiles(n, nb):
                if i0 < j0:
                    continue
                upd = (alpha * (A[i0:i1, k0:k1] @ B[j0:j1, k0:k1].T) + alpha * (B[i0:i1, k0:k1] @ A[j0:j1, k0:k1].T)
                       + b * C[i0:i1, j0:j1])
                if i0 == j0:
                    mask = np.tril(np.ones_like(upd, dtype=bool))
                    C[i0:i1, j0:j1][mask] = upd[mask]
                else:
                    C[i0:i1, j0:j1] = upd
    return C


def he_full(A_lower):
    """Full Hermitian matrix from its stored lower triangle (diagonal taken real)."""
    L = np.tril(A_lower)
    F = L + np.tril(L, -1).conj().T
    if np.iscomplexobj(F):
        d = np.arange(F.shape[0])
        F[d, d] = F[d, d].real
    return F


def hemm(alpha, A_lower, B, beta, C, nb: int, side: str = "L"):
    """C = alpha A B + beta C (Side::Left) or C = alpha B A + beta C (Side::Right), A Hermitian given by its lower
    triangle; accumulated one block column (Left) / block row (Right) of A at a time (src/hemmC.cc, src/hemmA.cc;
    Side::Right is the Left algorithm on conjugate-transposed views, src/hemmC.cc:57-70)."""
    A = he_full(A_lower)
    C = beta * np.array(C, order="F", copy=True)
    for (k0, k1) in _tiles(A.shape[0], nb):
        if side == "L":
            C += alpha * (A[:, k0:k1] @ B[k0:k1])
        else:
            C += alpha * (B[:, k0:k1] @ A[k0:k1, :])
    return C


def sy_full(A_lower):
    """Full (complex-)symmetric matrix from its stored lower triangle (no conjugation, diagonal as stored)."""
    L = np.tril(A_lower)
    return L + np.tril(L, -1).T


def symm(alpha, A_lower, B, beta, C, nb: int, side: str = "L"):
    """C = alpha A B + beta C (Side::Left) or C = alpha B A + beta C (Side::Right), A symmetric given by its lower
    triangle; one block column / block row of A per step (src/symm.cc)."""
    A = sy_full(A_lower)
    C = beta * np.array(C, order="F", copy=True)
    for (k0, k1) in _tiles(A.shape[0], nb):
        if side == "L":
            C += alpha * (A[:, k0:k1] @ B[k0:k1])
        else:
            C += alpha * (B[:, k0:k1] @ A[k0:k1, :])
    return C


def trmm(alpha, A_lower, B, nb: int, unit: bool = False, side: str = "L", op: str = "N"):
    """B <- alpha op(A) B (Side::Left) or B <- alpha B op(A) (Side::Right), A lower triangular, op = N / T / C a
    transposed view of it (src/trmm.cc -> work::trmm, src/work/work_trmm.cc; Side::Right runs the Left algorithm on
    the (conjugate-)transposed views, :78-98).  Every block row (Left) / block column (Right) of the result is formed
    from the ORIGINAL B: the diagonal-tile product first, then the off-diagonal tiles in ascending order."""
    A = np.tril(A_lower).astype(np.result_type(A_lower, B), copy=True)
    if unit:
        np.fill_diagonal(A, 1.0)
    M = {"N": A, "T": A.T, "C": A.conj().T}[op]             # lower triangular for N, upper for T / C
    B0 = np.array(B, order="F", copy=True)
    out = np.empty_like(B0)
    tiles = _tiles(A.shape[0], nb)
    for (i0, i1) in tiles:
        if side == "L":
            acc = M[i0:i1, i0:i1] @ B0[i0:i1]
            for (k0, k1) in tiles:
                if k0 != i0 and np.any(M[i0:i1, k0:k1]):
                    acc += M[i0:i1, k0:k1] @ B0[k0:k1]
            out[i0:i1] = alpha * acc
        else:
            acc = B0[:, i0:i1] @ M[i0:i1, i0:i1]
            for (k0, k1) in tiles:
                if k0 != i0 and np.any(M[k0:k1, i0:i1]):
                    acc += B0[:, k0:k1] @ M[k0:k1, i0:i1]
            out[:, i0:i1] = alpha * acc
    return out


def norm_inf(A, hermitian_lower: bool = False):
    F = he_full(A) if hermitian_lower else np.asarray(A)
    return float(np.abs(F).sum(axis=1).max())


def iter_ref_converged(colnorms_R, colnorms_X, cte) -> bool:       # src/internal/internal_util.hh:121-138
    return not np.any(np.asarray(colnorms_R) > np.asarray(colnorms_X) * cte)


def solve_mixed(A, B, nb: int, hermitian: bool, itermax: int = 30, tol=None, use_fallback: bool = True,
                ib: int = 16):
    """posv_mixed / gesv_mixed <double, float> (src/posv_mixed.cc:111-297, src/gesv_mixed.cc:106-300).
    A: full matrix (gesv) or lower triangle (posv).  Returns (X, iter, info)."""
    # <double, float> or <complex<double>, complex<float>> (the explicit instantiations, gesv_mixed.cc:303-316)
    hi, lo = (np.complex128, np.complex64) if np.iscomplexobj(A) else (np.float64, np.float32)
    A = np.asarray(A, dtype=hi)
    B = np.asarray(B, dtype=hi)
    n = A.shape[0]
    eps = np.finfo(np.float64).eps
    tol = eps * np.sqrt(n) if tol is None else tol
    Afull = he_full(A) if hermitian else A
    cte = norm_inf(A, hermitian) * tol
    A_lo = (np.tril(A) if hermitian else A).astype(lo)
    if hermitian:
        F_lo, info = potrf(he_full(A_lo), nb)
        solve_lo = lambda R: potrs(F_lo, R.astype(lo), nb)
    else:
        F_lo, piv, info = getrf(A_lo, nb, ib)
        solve_lo = lambda R: getrs(F_lo, piv, R.astype(lo), nb)
    converged, it = False, 0
    X = np.zeros_like(B)
    if info != 0:
        it = -3
    else:
        X = solve_lo(B).astype(hi)
        R = B - Afull @ X
        cm = lambda M: np.abs(M).max(axis=0)
        if iter_ref_converged(cm(R), cm(X), cte):
            converged = True
        for iiter in range(itermax):
            if converged:
                break
            X = X + solve_lo(R).astype(hi)
            R = B - Afull @ X
            if iter_ref_converged(cm(R), cm(X), cte):
                it, converged = iiter + 1, True
    if not converged:
        if info == 0:
            it = -itermax - 1
        if use_fallback:
            if hermitian:
                F, info = potrf(Afull, nb)
                X = potrs(F, B, nb) if info == 0 else X
            else:
                F, piv, info = getrf(A, nb, ib)
                X = getrs(F, piv, B, nb) if info == 0 else X
    return X, it, info


# ----------------------------------------------------------------------------
# memory-bound tile kernels (src/cuda/*.cu; include/slate/internal/device.hh:92-281)
# ----------------------------------------------------------------------------
def geadd(alpha, A, beta, B):            # device_geadd.cu:61-85
    return alpha * A + beta * B


def gescale(numer, denom, A):            # device_gescale.cu:42-60   (A *= numer/denom)
    return A * (numer / denom)


def gescale_row_col(equed: str, R, C, A):  # device_gescale_row_col.cu:42-140
    out = np.array(A, copy=True)
    if equed in ("R", "B"):
        out = out * np.asarray(R)[:, None]
    if equed in ("C", "B"):
        out = out * np.asarray(C)[None, :]
    return out


def geset(offdiag, diag, m, n, dtype=np.float64):   # device_geset.cu:44-70
    A = np.full((m, n), offdiag, dtype=dtype, order="F")
    d = np.arange(min(m, n))
    A[d, d] = diag
    return A


def tz_mask(uplo: str, m: int, n: int):
    i = np.arange(m)[:, None]
    j = np.arange(n)[None, :]
    return (i >= j) if uplo == "L" else (i <= j)


def tzset(uplo, offdiag, diag, A):       # device_tzset.cu:23-75
    out = np.array(A, copy=True)
    msk = tz_mask(uplo, *A.shape)
    out[msk] = offdiag
    d = np.arange(min(A.shape))
    out[d, d] = diag
    return out


def tzadd(uplo, alpha, A, beta, B):      # device_tzadd.cu:43-70
    out = np.array(B, copy=True)
    msk = tz_mask(uplo, *A.shape)
    out[msk] = (alpha * A + beta * B)[msk]
    return out


def tzscale(uplo, numer, denom, A):      # device_tzscale.cu:42-66
    out = np.array(A, copy=True)
    msk = tz_mask(uplo, *A.shape)
    out[msk] = (A * (numer / denom))[msk]
    return out


def tzcopy(uplo, A, B, dtype):           # device_tzcopy.cu:41-68
    out = np.array(B, copy=True).astype(dtype)
    msk = tz_mask(uplo, *A.shape)
    out[msk] = A.astype(dtype)[msk]
    return out


def _max_nan(a):
    """NaN-propagating max (device_util.cuh:22-25)."""
    a = np.asarray(a)
    return np.nan if np.isnan(a).any() else (a.max() if a.size else 0.0)


def genorm(norm: str, A):
    """Per-tile partial result of device::genorm (device_genorm.cu:44-281):
    'M' -> scalar max|a|; 'O' -> column sums (n); 'I' -> row sums (m);
    'F' -> (scale, sumsq) with scale^2 * sumsq = sum |a|^2."""
    absA = np.abs(A)
    if norm == "M":
        return _max_nan(absA)
    if norm == "O":
        return absA.sum(axis=0)
    if norm == "I":
        return absA.sum(axis=1)
    if norm == "F":
        scale = _max_nan(absA) if A.size else 0.0
        if scale == 0 or np.isnan(scale):
            return np.array([scale if np.isnan(scale) else 0.0, 1.0])
        return np.array([scale, ((absA / scale) ** 2).sum()])
    raise ValueError(norm)


def genorm_colmax(A):                    # ge_col_norms_max_kernel (device_genorm.cu:285-)
    return np.array([_max_nan(np.abs(A[:, j])) for j in range(A.shape[1])])


def henorm(norm: str, uplo: str, A):
    """Hermitian tile stored in one triangle (device_henorm.cu): result of the full matrix."""
    msk = tz_mask(uplo, *A.shape)
    T = np.where(msk, A, 0)
    F = T + np.conj(np.where(msk & ~np.eye(A.shape[0], dtype=bool), A, 0)).T
    if np.iscomplexobj(F):
        d = np.arange(F.shape[0])
        F[d, d] = F[d, d].real
    if norm == "M":
        return _max_nan(np.abs(F))
    if norm in ("O", "I"):
        return np.abs(F).sum(axis=0)
    return genorm("F", F)


def trnorm(norm: str, uplo: str, diag: str, A):   # device_trnorm.cu
    msk = tz_mask(uplo, *A.shape)
    T = np.where(msk, A, 0).astype(A.dtype)
    if diag == "U":
        d = np.arange(min(A.shape))
        T[d, d] = 1
    return genorm(norm, T)


def combine_norm(norm: str, parts):
    """Reduce per-tile partials to the matrix norm the way src/norm.cc / internal_genorm.cc do."""
    if norm == "M":
        return _max_nan(np.array(parts))
    if norm == "F":
        scale, sumsq = 0.0, 1.0
        for (s, q) in parts:                      # combine_sumsq (device_util.cuh:243-262)
            if s > scale:
                sumsq = q + sumsq * (scale / s) ** 2 if s != 0 else sumsq
                scale = s
            elif scale != 0:
                sumsq = sumsq + q * (s / scale) ** 2
        return scale * np.sqrt(sumsq)
    raise ValueError(norm)


# ----------------------------------------------------------------------------
# tester residual checks (the acceptance criteria)
# ----------------------------------------------------------------------------
def gemm_check(alpha, A, B, beta, C0, C, seed: int = 7):
    """|| C X - (alpha A (B X) + beta C0 X) ||_1 / ||Y||_1  must be <= 3 eps
    (test/test_gemm.cc:137-157,192-208); X = rand n-by-nrhs (nrhs = 10)."""
    n = C.shape[1]
    X = generate("rand", n, 10, seed, C.dtype)
    Y = alpha * (A @ (B @ X)) + beta * (C0 @ X)
    y_norm = np.abs(Y).sum(axis=0).max()
    R = C @ X - Y
    return np.abs(R).sum(axis=0).max() / y_norm


def solve_residual(A, X, B):
    """|| B - A X ||_1 / (n ||A||_1 ||X||_1) <= tol * eps / 2, tol = 50
    (test/test_posv.cc:304-345, test/test_gesv.cc:330-380, test/test.cc:356)."""
    n = A.shape[0]
    R = B - A @ X
    one = lambda M: np.abs(M).sum(axis=0).max()
    return one(R) / (n * one(A) * one(X))


def flops_gemm(m, n, k, complex_=False):      # blaspp/include/blas/flops.hh:100-104,312-324
    return (8.0 if complex_ else 2.0) * m * n * k


def flops_potrf(n):                           # lapackpp/include/lapack/flops.hh:56-60
    n = float(n)
    return (n ** 3 / 6 + n ** 2 / 2 + n / 3) + (n ** 3 / 6 - n / 6)


def flops_getrf(m, n):                        # lapackpp/include/lapack/flops.hh:27-39 (m >= n)
    m, n = float(m), float(n)
    if m >= n:
        fm = m * n * n / 2 - n ** 3 / 6 + m * n / 2 - n * n / 2 + 2 * n / 3
        fa = m * n * n / 2 - n ** 3 / 6 - m * n / 2 + n / 6
    else:
        fm = n * m * m / 2 - m ** 3 / 6 + n * m / 2 - m * m / 2 + 2 * m / 3
        fa = n * m * m / 2 - m ** 3 / 6 - n * m / 2 + m / 6
    return fm + fa
