"""Host-side mirror of the reference's matrix classes and drivers for the hot path.

Names, argument meaning and error behaviour follow the reference's C++ API:
  slate::Matrix / slate::HermitianMatrix   (include/slate/Matrix.hh, HermitianMatrix.hh)
  slate::multiply / gemm                   (include/slate/simplified_api.hh, src/gemm.cc)
  slate::chol_factor / potrf               (src/potrf.cc)     -> returns LAPACK-style info
  slate::lu_factor / getrf                 (src/getrf.cc)     -> returns info, fills Pivots
Everything computes on the GPU through libslate_b200.so (C ABI, include/slate_b200.h);
torch is used only for pinned host buffers, streams and torch.distributed bootstrap.
There is no CPU fallback: without the CUDA library/device every call raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import lib, check, SB200Error, c_i64, c_int, c_dbl, c_ptr, _sig


class Exception_(SB200Error):
    """slate::Exception equivalent (include/slate/Exception.hh)."""


class _Options(ctypes.Structure):
    _fields_ = [("lookahead", c_i64), ("inner_blocking", c_i64), ("pivot_threshold", c_dbl),
                ("reserved", c_int * 8)]


_grid_unique_id = _sig("sb200_grid_unique_id", [c_ptr])
_grid_create = _sig("sb200_grid_create", [c_int, c_int, c_int, c_ptr, ctypes.POINTER(c_ptr)])
_grid_destroy = _sig("sb200_grid_destroy", [c_ptr])
_matrix_create = _sig("sb200_matrix_create_d", [c_ptr, c_int, c_int, c_i64, c_i64, c_i64, ctypes.POINTER(c_ptr)])
_matrix_destroy = _sig("sb200_matrix_destroy", [c_ptr])
_matrix_generate = _sig("sb200_matrix_generate_d", [c_ptr, c_int, c_i64, c_ptr])
_matrix_from_host = _sig("sb200_matrix_from_host_d", [c_ptr, c_ptr, c_i64, c_ptr])
_matrix_to_host = _sig("sb200_matrix_to_host_d", [c_ptr, c_ptr, c_i64, c_ptr])
_matrix_from_host_local = _sig("sb200_matrix_from_host_local_d", [c_ptr, c_ptr, c_ptr])
_matrix_to_host_local = _sig("sb200_matrix_to_host_local_d", [c_ptr, c_ptr, c_ptr])
_last_panel_ms = _sig("sb200_last_driver_panel_ms", [c_ptr], c_dbl)
_matrix_copy = _sig("sb200_matrix_copy_d", [c_ptr, c_ptr, c_ptr])
_matrix_local_tiles = _sig("sb200_matrix_local_tiles", [c_ptr], c_i64)
_last_ms = _sig("sb200_last_driver_ms", [c_ptr], c_dbl)
_potrf = _sig("sb200_potrf_d", [c_ptr, ctypes.POINTER(_Options), ctypes.POINTER(c_i64)])
_gemm = _sig("sb200_gemm_d", [c_dbl, c_ptr, c_ptr, c_dbl, c_ptr, ctypes.POINTER(_Options)])
_getrf = _sig("sb200_getrf_d", [c_ptr, ctypes.POINTER(c_i64), ctypes.POINTER(_Options), ctypes.POINTER(c_i64)])


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _opts(opts):
    o = _Options()
    o.lookahead = int((opts or {}).get("lookahead", 1))
    o.inner_blocking = int((opts or {}).get("inner_blocking", 16))
    o.pivot_threshold = float((opts or {}).get("pivot_threshold", 1.0))
    return o


class Grid:
    """p x q process grid, one process per GPU, column-major rank order
    (slate::GridOrder::Col: rank = (i % p) + (j % q) * p, include/slate/func.hh:96-104).

    For p*q > 1 the NCCL communicator is bootstrapped over torch.distributed (any backend):
    rank 0 creates the ncclUniqueId and broadcasts it."""

    def __init__(self, p: int = 1, q: int = 1, rank: int = 0, unique_id: bytes | None = None):
        self.p, self.q, self.rank = p, q, rank
        h = c_ptr()
        buf = ctypes.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(_grid_create(p, q, rank, buf, ctypes.byref(h)), "Grid")
        self._h = h

    @staticmethod
    def choose(nranks: int):
        """As square as possible with p <= q (reference tester: test/test.cc:738-747)."""
        p = int(np.floor(np.sqrt(nranks)))
        while nranks % p:
            p -= 1
        return p, nranks // p

    @classmethod
    def from_torch_distributed(cls, p: int | None = None, q: int | None = None):
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        if p is None or q is None:
            p, q = cls.choose(world)
        if p * q != world:
            raise Exception_(f"grid {p}x{q} does not match world size {world}")
        if world == 1:
            return cls(1, 1, 0)
        uid = ctypes.create_string_buffer(128)
        if rank == 0:
            check(_grid_unique_id(uid), "ncclGetUniqueId")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0)
        return cls(p, q, rank, bytes(t.cpu().numpy().tobytes()))

    def close(self):
        if getattr(self, "_h", None):
            _grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_GRID: Grid | None = None


def default_grid() -> Grid:
    """The 1 x 1 grid (this process's current GPU) shared by matrices created without a grid,
    so that operands of one call live on the same grid object."""
    global _DEFAULT_GRID
    if _DEFAULT_GRID is None or _DEFAULT_GRID._h is None:
        _DEFAULT_GRID = Grid()
    return _DEFAULT_GRID


class Matrix:
    """General m-by-n tile matrix, nb-by-nb tiles, 2-D block-cyclic over the grid, resident in HBM."""
    _kind = "G"

    def __init__(self, m: int, n: int, nb: int, grid: Grid | None = None):
        self.grid = grid or default_grid()
        self.m, self.n, self.nb = int(m), int(n), int(nb)
        h = c_ptr()
        check(_matrix_create(self.grid._h, ord(self._kind), ord("C"), self.m, self.n, self.nb, ctypes.byref(h)),
              "Matrix")
        self._h = h

    # -- data movement -------------------------------------------------------------------
    def generate(self, kind: str = "rand", seed: int = 42):
        """Fill with the reference's test matrix (Philox-2x64 on global indices;
        matgen/generate_type_rand.hh): 'rand' or 'rand_dominant'."""
        code = {"rand": 0, "rand_dominant": 1}[kind]
        check(_matrix_generate(self._h, code, int(seed), _stream()), "generate")
        return self

    def from_host(self, hA, sync: bool = True):
        """Copy locally-owned tiles from a host column-major (m, n) array of the GLOBAL matrix.
        `hA`: numpy F-ordered float64 array or a torch CPU tensor whose memory is column-major
        (i.e. a (n, m) row-major tensor); pinned memory makes the copy asynchronous."""
        ptr, lda = _host_ptr(hA, self.m, self.n)
        check(_matrix_from_host(self._h, ptr, lda, _stream()), "from_host")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return self

    def to_host(self, out=None):
        import torch
        if out is None:
            out = np.zeros((self.m, self.n), dtype=np.float64, order="F")
        ptr, lda = _host_ptr(out, self.m, self.n)
        check(_matrix_to_host(self._h, ptr, lda, _stream()), "to_host")
        torch.cuda.current_stream().synchronize()
        return out

    def from_host_local(self, htiles, sync: bool = True):
        """Copy this rank's tiles from a packed host buffer (torch CPU float64 tensor, ideally pinned,
        of local_tiles * nb * nb elements in pool order: local block column, then local block row)."""
        self._check_local(htiles)
        check(_matrix_from_host_local(self._h, htiles.data_ptr(), _stream()), "from_host_local")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return self

    def to_host_local(self, htiles, sync: bool = True):
        self._check_local(htiles)
        check(_matrix_to_host_local(self._h, htiles.data_ptr(), _stream()), "to_host_local")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return htiles

    def _check_local(self, t):
        import torch
        need = self.local_tiles * self.nb * self.nb
        if not (isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.dtype == torch.float64
                and t.is_contiguous() and t.numel() == need):
            raise Exception_(f"local tile buffer must be a contiguous float64 CPU tensor of {need} elements")

    def copy_from(self, other: "Matrix"):
        check(_matrix_copy(self._h, other._h, _stream()), "copy")
        return self

    @property
    def local_tiles(self) -> int:
        return int(_matrix_local_tiles(self._h))

    @property
    def last_driver_ms(self) -> float:
        return float(_last_ms(self._h))

    @property
    def last_panel_ms(self) -> float:
        return float(_last_panel_ms(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HermitianMatrix(Matrix):
    """n-by-n Hermitian (real: symmetric) matrix, lower tiles stored
    (slate::HermitianMatrix with Uplo::Lower, include/slate/HermitianMatrix.hh)."""
    _kind = "H"

    def __init__(self, n: int, nb: int, grid: Grid | None = None, uplo: str = "L"):
        if uplo.upper()[0] != "L":
            raise Exception_("only Uplo::Lower storage is implemented")
        super().__init__(n, n, nb, grid)


def _host_ptr(a, m, n):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            if a.device.type != "cpu" or a.dtype != torch.float64 or not a.is_contiguous():
                raise Exception_("host tensor must be a contiguous float64 CPU tensor")
            if tuple(a.shape) != (n, m):
                raise Exception_(f"host tensor holding a column-major {m}x{n} matrix must have shape ({n}, {m})")
            return a.data_ptr(), max(m, 1)
    except ImportError:
        pass
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.shape != (m, n) or not a.flags.f_contiguous:
        raise Exception_(f"host array must be float64, shape ({m}, {n}), Fortran order")
    return a.ctypes.data, max(m, 1)


# -- drivers ---------------------------------------------------------------------------------
def gemm(alpha: float, A: Matrix, B: Matrix, beta: float, C: Matrix, opts: dict | None = None):
    """C = alpha A B + beta C  (slate::gemm, src/gemm.cc:82-105 -> gemmC)."""
    o = _opts(opts)
    check(_gemm(float(alpha), A._h, B._h, float(beta), C._h, ctypes.byref(o)), "gemm")


multiply = gemm     # simplified API name (include/slate/simplified_api.hh)


def potrf(A: HermitianMatrix, opts: dict | None = None) -> int:
    """Cholesky A = L L^H, lower (slate::potrf, src/potrf.cc:262-281).
    Returns info: 0, or i > 0 if the leading minor of order i is not positive definite."""
    o = _opts(opts)
    info = c_i64(0)
    check(_potrf(A._h, ctypes.byref(o), ctypes.byref(info)), "potrf")
    return int(info.value)


chol_factor = potrf


def getrf(A: Matrix, opts: dict | None = None):
    """LU with partial pivoting P A = L U (slate::getrf, src/getrf.cc:320-350).
    Returns (pivots, info); pivots[k] = list of (tileIndex, elementOffset) relative to the
    panel sub-matrix A(k:mt-1, k), as slate::Pivots (include/slate/types.hh:84-117)."""
    o = _opts(opts)
    info = c_i64(0)
    mn = min(A.m, A.n)
    flat = (c_i64 * (2 * max(mn, 1)))()
    check(_getrf(A._h, flat, ctypes.byref(o), ctypes.byref(info)), "getrf")
    piv = np.frombuffer(flat, dtype=np.int64)[: 2 * mn].reshape(-1, 2)
    nb = A.nb
    pivots = [[(int(t), int(off)) for t, off in piv[k0:min(k0 + nb, mn)]] for k0 in range(0, mn, nb)]
    return pivots, int(info.value)


lu_factor = getrf
