"""CPU check of the LOGIC of tests/test_zzzzz_gpu_blas3_variants.py (shapes, argument conventions, golden files, tolerances)
-- NOT of the product.  The GPU test functions are called with a numpy stand-in for `slate_b200.host` that computes every
routine by its textbook formula in the matrix's own precision (float32 work for s / c), so a test that would fail on a
correct implementation -- wrong shape, wrong golden parameters, a tolerance below rounding -- shows up here instead of as
an XFAIL on the GPU box that would be blamed on the library.  Test infrastructure only (scratch/): nothing in the product
or in the test suite imports this file.

    python scratch/cpu_standin/check_gpu_test_logic.py [-k substring]
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import slate_oracle as o                                    # noqa: E402  (generator + LU pivots only)

NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
OP = {"N": lambda x: x, "T": lambda x: x.T, "C": lambda x: x.conj().T}


class SB200Error(Exception):
    pass


class Matrix:
    _kind = "G"

    def __init__(self, m, n, nb, grid=None, dtype="d"):
        self.t = dtype if isinstance(dtype, str) else {np.float32: "s", np.float64: "d", np.complex64: "c", np.complex128: "z"}[dtype]
        self.dtype = np.dtype(NP[self.t])
        self.m, self.n, self.nb = m, n, nb
        self.a = np.zeros((m, n), dtype=self.dtype, order="F")

    def generate(self, kind, seed):
        self.a = np.asfortranarray(o.generate(kind, self.m, self.n, seed, self.dtype))
        return self

    def from_host(self, h):
        self.a = np.array(h, dtype=self.dtype, order="F")
        return self

    def to_host(self):
        return self.a.copy()


class HermitianMatrix(Matrix):
    _kind = "H"

    def __init__(self, n, nb, grid=None, uplo="L", dtype="d"):
        super().__init__(n, n, nb, grid, dtype)

    def lower(self):
        return np.tril(self.a)

    def he_full(self):
        L = self.lower()
        F = L + np.tril(L, -1).conj().T
        if np.iscomplexobj(F):
            F[np.diag_indices_from(F)] = F.diagonal().real
        return F

    def sy_full(self):
        L = self.lower()
        return L + np.tril(L, -1).T


def _tri(A, uplo, diag):
    T = np.tril(A.a) if uplo == "L" else np.triu(A.a)
    if diag == "U":
        T = T.copy(); np.fill_diagonal(T, 1)
    return T


def _sc(M, v):
    return M.dtype.type(v) if not np.iscomplexobj(M.a) else M.dtype.type(complex(v))


def _chk(cond):
    if not cond:
        raise SB200Error("invalid argument")


def gemm(alpha, A, B, beta, C, opts=None, opA="N", opB="N"):
    a, b = OP[opA](A.a), OP[opB](B.a)
    _chk(a.shape[0] == C.m and b.shape[1] == C.n and a.shape[1] == b.shape[0])
    C.a = np.asfortranarray(_sc(C, alpha) * (a @ b) + _sc(C, beta) * C.a)


def _store_lower(C, full):
    C.a = np.asfortranarray(np.tril(full) + np.triu(C.a, 1))


def herk(alpha, A, beta, C, opts=None, op="N"):
    if op != "N":
        _chk(not (np.iscomplexobj(A.a) and op != "C"))
    a = A.a if op == "N" else A.a.conj().T
    _chk(a.shape[0] == C.n)
    full = C.dtype.type(alpha) * (a @ a.conj().T) + C.dtype.type(beta) * C.he_full() if not np.iscomplexobj(C.a) else \
        np.float64(alpha).astype(C.a.real.dtype) * (a @ a.conj().T) + np.float64(beta).astype(C.a.real.dtype) * _he_keep_diag(C)
    if np.iscomplexobj(full):
        full[np.diag_indices_from(full)] = full.diagonal().real
    _store_lower(C, full)


def _he_keep_diag(C):
    """lower triangle mirrored; the stored (possibly complex) diagonal enters beta * C as it is, as in the tile kernels"""
    L = C.lower()
    return L + np.tril(L, -1).conj().T


def her2k(alpha, A, B, beta, C, opts=None, op="N"):
    if op != "N":
        _chk(not (np.iscomplexobj(A.a) and op != "C"))
    a, b = (A.a, B.a) if op == "N" else (A.a.conj().T, B.a.conj().T)
    al = _sc(C, alpha)
    full = al * (a @ b.conj().T) + np.conj(al) * (b @ a.conj().T) + C.a.real.dtype.type(beta) * _he_keep_diag(C)
    if np.iscomplexobj(full):
        full[np.diag_indices_from(full)] = full.diagonal().real
    _store_lower(C, full)


def syrk(alpha, A, beta, C, opts=None, op="N"):
    if op != "N":
        _chk(not (np.iscomplexobj(A.a) and op != "T"))
    a = A.a if op == "N" else A.a.T
    _store_lower(C, _sc(C, alpha) * (a @ a.T) + _sc(C, beta) * C.sy_full())


def syr2k(alpha, A, B, beta, C, opts=None, op="N"):
    if op != "N":
        _chk(not (np.iscomplexobj(A.a) and op != "T"))
    a, b = (A.a, B.a) if op == "N" else (A.a.T, B.a.T)
    _store_lower(C, _sc(C, alpha) * (a @ b.T + b @ a.T) + _sc(C, beta) * C.sy_full())


def hemm(alpha, A, B, beta, C, opts=None, side="L"):
    F = A.he_full()
    _chk((B.m if side == "L" else B.n) == A.n and B.a.shape == C.a.shape)
    C.a = np.asfortranarray(_sc(C, alpha) * (F @ B.a if side == "L" else B.a @ F) + _sc(C, beta) * C.a)


def symm(alpha, A, B, beta, C, opts=None, side="L"):
    F = A.sy_full()
    _chk((B.m if side == "L" else B.n) == A.n and B.a.shape == C.a.shape)
    C.a = np.asfortranarray(_sc(C, alpha) * (F @ B.a if side == "L" else B.a @ F) + _sc(C, beta) * C.a)


def trmm(alpha, A, B, side="L", uplo="L", op="N", diag="N", opts=None):
    _chk(uplo == "L")
    _chk((B.m if side == "L" else B.n) == A.n)
    M = OP[op](_tri(A, "L", diag))
    B.a = np.asfortranarray(_sc(B, alpha) * (M @ B.a if side == "L" else B.a @ M))


def trsm(alpha, A, B, side="L", uplo="L", op="N", diag="N", opts=None):
    _chk(not (A._kind == "H" and uplo != "L"))
    _chk((B.m if side == "L" else B.n) == A.n)
    B.a = np.asfortranarray(o.trsm_tile(side, uplo, op, diag, _sc(B, alpha), A.a, B.a).astype(B.dtype))


triangular_solve = trsm


def potrf(A, opts=None):
    F = A.he_full()
    try:
        L = np.linalg.cholesky(F.astype(np.complex128 if np.iscomplexobj(F) else np.float64))
    except np.linalg.LinAlgError:
        _, info = o.potrf(F, A.nb)
        return info
    A.a = np.asfortranarray(L.astype(A.dtype))
    return 0


def potrs(A, B, opts=None):
    L = np.tril(A.a)
    B.a = np.asfortranarray(np.linalg.solve(L.conj().T, np.linalg.solve(L, B.a)).astype(B.dtype))


def posv(A, B, opts=None):
    info = potrf(A, opts)
    if info == 0:
        potrs(A, B, opts)
    return info


chol_solve = posv


def getrf(A, opts=None):
    LU, piv, info = o.getrf(A.a.astype(np.complex128 if np.iscomplexobj(A.a) else np.float64), A.nb)
    A.a = np.asfortranarray(LU.astype(A.dtype))
    return piv, info


def getrs(A, pivots, B, opts=None, op="N"):
    n = A.n
    perm = o.pivots_to_perm(pivots, n, A.nb)
    L = np.tril(A.a, -1) + np.eye(n, dtype=A.dtype)
    U = np.triu(A.a)
    P = np.eye(n)[perm]
    M = OP[op](P.T @ (L @ U))                      # A = P^T L U
    B.a = np.asfortranarray(np.linalg.solve(M, B.a).astype(B.dtype))


def gesv(A, B, opts=None):
    piv, info = getrf(A, opts)
    if info == 0:
        getrs(A, piv, B, opts)
    return piv, info


lu_solve = gesv


def getrf_nopiv(A, opts=None):
    LU, info = o.getrf_nopiv(A.a.astype(np.float64), A.nb)
    A.a = np.asfortranarray(LU.astype(A.dtype))
    return info


def lu_solve_nopiv(A, B, opts=None):
    a0 = A.a.copy()
    info = getrf_nopiv(A, opts)
    if info == 0:
        B.a = np.asfortranarray(np.linalg.solve(a0.astype(np.float64), B.a.astype(np.float64)).astype(B.dtype))
    return info


def norm(kind, A, symmetric=False):
    if kind not in ("max", "one", "inf", "fro"):
        raise SB200Error("unknown norm")
    F = A.a if A._kind == "G" else (A.sy_full() if symmetric else A.he_full())
    a = np.abs(F.astype(np.complex128 if np.iscomplexobj(F) else np.float64))
    if np.isnan(a).any():
        return float("nan")
    return float({"max": a.max(), "one": a.sum(axis=0).max(), "inf": a.sum(axis=1).max(), "fro": np.sqrt((a * a).sum())}[kind])


def norm_inf(A):
    return norm("inf", A)


def main():
    import types
    import pytest
    sl = types.SimpleNamespace(**{k: v for k, v in globals().items() if not k.startswith("_")})
    sys.modules.setdefault("torch", types.ModuleType("torch"))           # the one `import torch` inside a test body
    import importlib
    mod = importlib.import_module("tests.test_zzzzz_gpu_blas3_variants")
    golden = os.path.join(ROOT, "tests", "golden")
    sel = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "-k" else ""
    ran = failed = 0
    for name in sorted(n for n in dir(mod) if n.startswith("test_")):
        f = getattr(mod, name)
        if sel and sel not in name:
            continue
        if "reference_tester" in name:
            continue                                                    # subprocesses of the GPU tester: not logic of this kind
        marks = [m for m in getattr(f, "pytestmark", []) if m.name == "parametrize"]
        axes = []
        for m in marks:
            names = [x.strip() for x in m.args[0].split(",")]
            axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in m.args[1]])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo:
                kw.update(d)
            argnames = f.__code__.co_varnames[:f.__code__.co_argcount]
            if "sl" in argnames:
                kw["sl"] = sl
            if "golden_dir" in argnames:
                kw["golden_dir"] = golden
            ran += 1
            try:
                f(**kw)
            except Exception as ex:   # noqa: BLE001
                failed += 1
                print(f"FAIL {name} {({k: v for k, v in kw.items() if k not in ('sl', 'golden_dir')})}: {type(ex).__name__}: {str(ex)[:200]}")
    print(f"{ran} cases of the GPU test file run against the numpy stand-in: {ran - failed} pass, {failed} fail")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
