"""CPU check of the SCHEDULES of the side / op variants of trmm / hemm / symm / trsm (slate_b200/csrc/solve.cu:
trmm_lower_variant, hemm_symm_right_lower, tri_sweep_right, gemm_ops, rank_update_trans, norm_mat): the step order, the batches of a step, the operand roles and the in-place
update through a one-block workspace are restated here tile by tile in numpy, exactly as the driver issues them, and
compared with the oracle (which is pinned to the unmodified reference's golden output, tests/test_oracle.py).  What this
does NOT cover is the C++ transcription and the kernels: that is tests/test_zzzzz_gpu_blas3_variants.py on a GPU."""
import numpy as np
import pytest

from oracle import slate_oracle as o

EPS = np.finfo(np.float64).eps
OP = {"N": lambda x: x, "T": lambda x: x.T, "C": lambda x: x.conj().T}


def tiles(n, nb):
    return [(i, min(i + nb, n)) for i in range(0, n, nb)]


def gemm(opa, opb, alpha, a, b, beta, c):
    """what one entry of a batched launch does: c <- alpha opa(a) opb(b) + beta c, c not aliased with a / b"""
    return alpha * (OP[opa](a) @ OP[opb](b)) + beta * c


def trmm_lower_variant(side, op, alpha, A_lower, B, nb, unit):
    A = np.array(A_lower)                      # only tiles (i, j) with i >= j are read, as with kind 'H' storage
    B = np.array(B, order="F", copy=True)
    left = side == "L"
    assert not (left and op == "N")
    ta = tiles(A.shape[0], nb)
    tr, tc = tiles(B.shape[0], nb), tiles(B.shape[1], nb)
    nt = len(ta)
    dtri = []
    for (k0, k1) in ta:                        # tr_fill_kernel
        d = np.tril(A[k0:k1, k0:k1]).copy()
        if unit:
            np.fill_diagonal(d, 1.0)
        dtri.append(d)
    opa, opb = (op, "N") if left else ("N", op)
    order = range(nt) if (left or op == "N") else range(nt - 1, -1, -1)
    for k in order:
        k0, k1 = ta[k]
        snapshot = B.copy()                    # a batched launch reads its operands while other entries write THEIR targets:
        if left:                               # sources and targets of one launch must be disjoint -- asserted below
            targets = [(i, j) for j in range(len(tc)) for i in range(k)]
            for (i, j) in targets:
                (i0, i1), (j0, j1) = tr[i], tc[j]
                assert i != k
                B[i0:i1, j0:j1] = gemm(opa, opb, alpha, A[k0:k1, i0:i1], snapshot[k0:k1, j0:j1], 1.0, snapshot[i0:i1, j0:j1])
            w = [gemm(opa, opb, alpha, dtri[k], B[k0:k1, j0:j1], 0.0, 0.0) for (j0, j1) in tc]
            for (j0, j1), wj in zip(tc, w):
                B[k0:k1, j0:j1] = wj
        else:
            js = range(0, k) if op == "N" else range(k + 1, nt)
            for i, (i0, i1) in enumerate(tr):
                for j in js:
                    j0, j1 = tc[j]
                    a_tile = A[k0:k1, j0:j1] if op == "N" else A[j0:j1, k0:k1]        # stored: row index >= column index
                    assert (k >= j) if op == "N" else (j >= k)
                    B[i0:i1, j0:j1] = gemm(opa, opb, alpha, snapshot[i0:i1, k0:k1], a_tile, 1.0, snapshot[i0:i1, j0:j1])
            w = [gemm(opa, opb, alpha, B[i0:i1, k0:k1], dtri[k], 0.0, 0.0) for (i0, i1) in tr]
            for (i0, i1), wi in zip(tr, w):
                B[i0:i1, k0:k1] = wi
    return B


def hemm_symm_right_lower(conj, alpha, A_lower, X, beta, C, nb):
    A = np.array(A_lower)
    C = beta * np.array(C, order="F", copy=True)           # scale_kernel
    ta = tiles(A.shape[0], nb)
    tr = tiles(X.shape[0], nb)
    oph = "C" if conj else "T"
    dfull = []
    for (k0, k1) in ta:                                    # he_fill_kernel / sy_fill_kernel
        d = np.tril(A[k0:k1, k0:k1])
        if conj:
            full = d + np.tril(d, -1).conj().T
            full[np.diag_indices_from(full)] = np.real(np.diag(d))
        else:
            full = d + np.tril(d, -1).T
        dfull.append(full)
    for k, (k0, k1) in enumerate(ta):
        for (i0, i1) in tr:
            for j, (j0, j1) in enumerate(ta):
                if j < k:
                    C[i0:i1, j0:j1] = gemm("N", "N", alpha, X[i0:i1, k0:k1], A[k0:k1, j0:j1], 1.0, C[i0:i1, j0:j1])
                elif j > k:
                    C[i0:i1, j0:j1] = gemm("N", oph, alpha, X[i0:i1, k0:k1], A[j0:j1, k0:k1], 1.0, C[i0:i1, j0:j1])
                else:
                    C[i0:i1, k0:k1] = gemm("N", "N", alpha, X[i0:i1, k0:k1], dfull[k], 1.0, C[i0:i1, k0:k1])
    return C


def tri_sweep_right(A_tri, lower, op, unit, B, nb):
    """solve.cu: tri_sweep_right -- per step one right-side diagonal-tile solve of block column k (trsm_colmajor, whose
    result the oracle's trsm_tile states) and one batched update of the block columns that still wait"""
    A = np.array(A_tri)
    B = np.array(B, order="F", copy=True)
    ta = tiles(A.shape[0], nb)
    tr = tiles(B.shape[0], nb)
    kt = len(ta)
    trans = op != "N"
    forward = not (lower != trans)
    for sidx in range(kt):
        k = sidx if forward else kt - 1 - sidx
        k0, k1 = ta[k]
        for (i0, i1) in tr:
            B[i0:i1, k0:k1] = o.trsm_tile("R", "L" if lower else "U", op, "U" if unit else "N", 1.0, A[k0:k1, k0:k1], B[i0:i1, k0:k1])
        js = range(k + 1, kt) if forward else range(0, k)
        snapshot = B.copy()
        for j in js:
            j0, j1 = ta[j]
            m_tile = A[j0:j1, k0:k1] if trans else A[k0:k1, j0:j1]
            assert np.isfinite(m_tile).all(), "the update read a tile outside the stored triangle"
            for (i0, i1) in tr:
                B[i0:i1, j0:j1] = gemm("N", op if trans else "N", -1.0, snapshot[i0:i1, k0:k1], m_tile, 1.0, snapshot[i0:i1, j0:j1])
    return B


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("lower,op", [(True, "N"), (True, "T"), (True, "C"), (False, "N"), (False, "T"), (False, "C")])
@pytest.mark.parametrize("unit", [False, True])
@pytest.mark.parametrize("m,n,nb", [(70, 200, 64), (128, 128, 64), (300, 50, 128), (10, 260, 128)])
def test_trsm_right_sweep_schedule_matches_oracle(dt, lower, op, unit, m, n, nb):
    G = o.generate("rand_dominant", n, n, 42, dt)
    tri = np.tril(G) if lower else np.triu(G)
    outside = np.triu(np.full((n, n), np.nan), 1) if lower else np.tril(np.full((n, n), np.nan), -1)
    # tiles wholly outside the triangle do not exist (lower storage) or are never to be read: poison them; the diagonal
    # tiles keep zeros outside the triangle as trsm_colmajor only reads their triangle
    poisoned = tri + outside
    for (k0, k1) in tiles(n, nb):
        poisoned[k0:k1, k0:k1] = tri[k0:k1, k0:k1]
    B = o.generate("rand", m, n, 43, dt)
    out = tri_sweep_right(poisoned, lower, op, unit, B, nb)
    ref = o.trsm(1.0, tri, B, nb, side="R", lower=lower, op=op, unit=unit)
    direct = o.trsm_tile("R", "L" if lower else "U", op, "U" if unit else "N", 1.0, tri, B)
    assert np.isfinite(out).all()
    scale = max(np.abs(ref).max(), 1.0)
    assert np.abs(out - ref).max() <= 1e-11 * scale and np.abs(direct - ref).max() <= 1e-11 * scale


def gemm_ops(opa, opb, alpha, A, B, beta, C, nb):
    """solve.cu: gemm_ops -- step k multiplies STORED tile (k, i) of A (opA != N) or (i, k), and STORED tile (j, k) of B
    (opB != N) or (k, j), with the op handed to the tile GEMM"""
    C = np.array(C, order="F", copy=True)
    ta, tb = opa != "N", opb != "N"
    tk = tiles(A.shape[0] if ta else A.shape[1], nb)
    for k, (k0, k1) in enumerate(tk):
        for (j0, j1) in tiles(C.shape[1], nb):
            for (i0, i1) in tiles(C.shape[0], nb):
                a_tile = A[k0:k1, i0:i1] if ta else A[i0:i1, k0:k1]
                b_tile = B[j0:j1, k0:k1] if tb else B[k0:k1, j0:j1]
                C[i0:i1, j0:j1] = gemm(opa, opb, alpha, a_tile, b_tile, beta if k == 0 else 1.0, C[i0:i1, j0:j1])
    return C


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("opa", ["N", "T", "C"])
@pytest.mark.parametrize("opb", ["N", "T", "C"])
@pytest.mark.parametrize("m,n,k,nb", [(150, 200, 100, 64), (70, 10, 300, 64)])
def test_gemm_ops_schedule_matches_oracle(dt, opa, opb, m, n, k, nb):
    A = o.generate("rand", *((m, k) if opa == "N" else (k, m)), 42, dt)
    B = o.generate("rand", *((k, n) if opb == "N" else (n, k)), 43, dt)
    C = o.generate("rand", m, n, 44, dt)
    al, be = (3.1 + 1.4j, 2.7 + 1.7j) if dt is np.complex128 else (3.1, 2.7)
    out = gemm_ops(opa, opb, al, A, B, be, C, nb)
    ref = al * (OP[opa](A) @ OP[opb](B)) + be * C
    assert np.abs(out - ref).max() <= 64 * EPS * np.abs(ref).max()
    assert np.abs(out - o.gemm(al, A, B, be, C, nb, opa=opa, opb=opb)).max() <= 64 * EPS * np.abs(ref).max()


def rank_update_trans(conj, alpha, A, B, beta, C_lower, nb):
    """solve.cu: rank_update_trans -- step kk multiplies STORED tiles (kk, i) and (kk, j) of the k x n matrices with the op on
    the A-role operand; diagonal tiles keep their lower triangle (epilogue mask), Hermitian diagonal forced real"""
    C = np.array(C_lower, order="F", copy=True)
    oph = "C" if conj else "T"
    tn = tiles(C.shape[0], nb)
    alpha2 = np.conj(alpha) if conj else alpha
    for kk, (k0, k1) in enumerate(tiles(A.shape[0], nb)):
        launches = [(A, B if B is not None else A, alpha, beta if kk == 0 else 1.0)]
        if B is not None:
            launches.append((B, A, alpha2, 1.0))
        for (X, Y, al, be) in launches:
            for j, (j0, j1) in enumerate(tn):
                for i, (i0, i1) in enumerate(tn):
                    if i < j:
                        continue
                    upd = gemm(oph, "N", al, X[k0:k1, i0:i1], Y[k0:k1, j0:j1], be, C[i0:i1, j0:j1])
                    if i == j:
                        mask = np.tril(np.ones(upd.shape, dtype=bool))
                        blk = C[i0:i1, j0:j1]
                        blk[mask] = upd[mask]
                        if conj and np.iscomplexobj(blk):
                            di = np.arange(blk.shape[0])
                            blk[di, di] = blk[di, di].real
                    else:
                        C[i0:i1, j0:j1] = upd
    return C


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("routine", ["herk", "her2k", "syrk", "syr2k"])
@pytest.mark.parametrize("n,k,nb", [(200, 100, 64), (128, 300, 64), (70, 10, 128)])
def test_rank_update_trans_schedule_matches_oracle(dt, routine, n, k, nb):
    A0 = o.generate("rand", k, n, 42, dt)
    B0 = o.generate("rand", k, n, 43, dt)
    C = np.tril(o.generate("rand", n, n, 44, dt))
    cplx = dt is np.complex128
    al, be = (ALPHA, BETA) if cplx else (ALPHA.real, BETA.real)
    if routine == "herk":
        out = rank_update_trans(True, al.real, A0, None, be.real, C, nb)
        ref = o.herk(al.real, A0.conj().T, be.real, C, nb)
    elif routine == "her2k":
        out = rank_update_trans(True, al, A0, B0, be.real, C, nb)
        ref = o.her2k(al, A0.conj().T, B0.conj().T, be.real, C, nb)
    elif routine == "syrk":
        out = rank_update_trans(False, al, A0, None, be, C, nb)
        ref = o.syrk(al, A0.T, be, C, nb)
    else:
        out = rank_update_trans(False, al, A0, B0, be, C, nb)
        ref = o.syr2k(al, A0.T, B0.T, be, C, nb)
    assert np.abs(np.tril(out) - np.tril(ref)).max() <= 64 * EPS * np.abs(np.tril(ref)).max()


def norm_mat(norm, A, nb, kind="G", symmetric=False):
    """solve.cu: norm_mat -- per-tile partial results (the oracle's restatements of the tile kernels) combined on the host:
    column / row sums per block column / row, an off-diagonal tile of a Hermitian / symmetric matrix also stands for its
    mirror image (row sums -> the mirror's columns, sumsq twice), (scale, sumsq) pairs as lassq"""
    m, n = A.shape
    tr, tc = tiles(m, nb), tiles(n, nb)
    he = kind == "H"

    def diag_tile(T):
        if symmetric:
            L = np.tril(T)
            return L + np.tril(L, -1).T
        return None

    parts_max, pairs = [], []
    acc = np.zeros(m if (not he and norm == "I") else n)
    for j, (j0, j1) in enumerate(tc):
        for i, (i0, i1) in enumerate(tr):
            if he and i < j:
                continue
            T = A[i0:i1, j0:j1]
            diag = he and i == j
            if diag:
                F = diag_tile(T)
                val = (lambda nm: o.genorm(nm, F)) if symmetric else (lambda nm: o.henorm(nm, "L", T))
            else:
                val = lambda nm: o.genorm(nm, T)
            if norm == "M":
                parts_max.append(val("M"))
            elif norm == "F":
                s, q = val("F")
                pairs.append((s, (2.0 if (he and not diag) else 1.0) * q))
            elif not he:
                if norm == "O":
                    acc[j0:j1] += val("O")
                else:
                    acc[i0:i1] += val("I")
            else:
                acc[j0:j1] += val("O")
                if not diag:
                    acc[i0:i1] += val("I")
    if norm == "M":
        return o.combine_norm("M", parts_max)
    if norm == "F":
        return o.combine_norm("F", pairs)
    return acc.max()


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("m,n,nb", [(200, 136, 64), (64, 64, 64), (50, 300, 128)])
def test_norm_combination_general(dt, m, n, nb):
    A = o.generate("rand", m, n, 42, dt)
    a = np.abs(A)
    ref = {"M": a.max(), "O": a.sum(axis=0).max(), "I": a.sum(axis=1).max(), "F": np.sqrt((a * a).sum())}
    for nm, r in ref.items():
        assert abs(norm_mat(nm, A, nb) - r) <= 64 * EPS * r, nm


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("symmetric", [False, True])
@pytest.mark.parametrize("n,nb", [(200, 64), (300, 512), (128, 64)])
def test_norm_combination_hermitian_and_symmetric(dt, symmetric, n, nb):
    L = np.tril(o.generate("rand", n, n, 42, dt))
    stored = L + np.triu(np.full((n, n), np.nan), 1)
    for (k0, k1) in tiles(n, nb):                       # diagonal tiles exist as whole tiles; their upper part is never read
        stored[k0:k1, k0:k1] = np.where(np.tril(np.ones((k1 - k0, k1 - k0), dtype=bool)), L[k0:k1, k0:k1], 7.0)
    full = o.sy_full(L) if symmetric else o.he_full(L)
    a = np.abs(full)
    ref = {"M": a.max(), "O": a.sum(axis=0).max(), "I": a.sum(axis=1).max(), "F": np.sqrt((a * a).sum())}
    for nm, r in ref.items():
        v = norm_mat(nm, stored, nb, kind="H", symmetric=symmetric)
        assert np.isfinite(v) and abs(v - r) <= 64 * EPS * r, nm


ALPHA = 3.141592653589793 + 1.414213562373095j
BETA = 2.718281828459045 + 1.732050807568877j


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("unit", [False, True])
@pytest.mark.parametrize("side,op", [("L", "T"), ("L", "C"), ("R", "N"), ("R", "T"), ("R", "C")])
@pytest.mark.parametrize("m,n,nb", [(200, 70, 64), (70, 200, 64), (128, 128, 64), (50, 300, 128)])
def test_trmm_variant_schedule_matches_oracle(dt, unit, side, op, m, n, nb):
    na = m if side == "L" else n
    A = np.tril(o.generate("rand", na, na, 42, dt))
    A = A + np.triu(np.full((na, na), np.nan), 1)          # the upper tiles do not exist in lower storage: reading them would show
    B = o.generate("rand", m, n, 43, dt)
    al = ALPHA if dt is np.complex128 else ALPHA.real
    out = trmm_lower_variant(side, op, al, A, B, nb, unit)
    ref = o.trmm(al, np.tril(np.nan_to_num(A)), B, nb, unit=unit, side=side, op=op)
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() <= 64 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("conj", [True, False])
@pytest.mark.parametrize("m,n,nb", [(70, 192, 64), (200, 200, 64), (10, 300, 128)])
def test_hemm_symm_right_schedule_matches_oracle(dt, conj, m, n, nb):
    A = np.tril(o.generate("rand", n, n, 42, dt))
    A = A + np.triu(np.full((n, n), np.nan), 1)
    X = o.generate("rand", m, n, 43, dt)
    C = o.generate("rand", m, n, 44, dt)
    al, be = (ALPHA, BETA) if dt is np.complex128 else (ALPHA.real, BETA.real)
    out = hemm_symm_right_lower(conj, al, A, X, be, C, nb)
    Al = np.tril(np.nan_to_num(A))
    ref = (o.hemm if conj else o.symm)(al, Al, X, be, C, nb, side="R")
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() <= 64 * EPS * np.abs(ref).max()
