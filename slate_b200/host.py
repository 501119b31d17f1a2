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

from ._lib import lib, check, SB200Error, c_i64, c_int, c_dbl, c_ptr, _sig, scalar, SCALAR_T, REAL_T

NP_DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
_DTYPE_CHAR = {np.dtype(v): k for k, v in NP_DTYPE.items()}


def dtype_char(dtype) -> str:
    """'s' | 'd' | 'c' | 'z' for a numpy dtype / type character."""
    if isinstance(dtype, str) and dtype in NP_DTYPE:
        return dtype
    try:
        return _DTYPE_CHAR[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise Exception_(f"unsupported element type {dtype!r}") from None


class Exception_(SB200Error):
    """slate::Exception equivalent (include/slate/Exception.hh)."""


class _Options(ctypes.Structure):
    _fields_ = [("lookahead", c_i64), ("inner_blocking", c_i64), ("pivot_threshold", c_dbl),
                ("reserved", c_int * 8)]


_grid_unique_id = _sig("sb200_grid_unique_id", [c_ptr])
_grid_create = _sig("sb200_grid_create", [c_int, c_int, c_int, c_ptr, ctypes.POINTER(c_ptr)])
_grid_destroy = _sig("sb200_grid_destroy", [c_ptr])
_bcast_tiles = _sig("sb200_bcast_tiles", [c_ptr, c_i64, ctypes.POINTER(c_ptr), ctypes.POINTER(c_ptr),
                                          ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(c_int), c_ptr])
class _MixedOptions(ctypes.Structure):
    _fields_ = [("max_iterations", c_i64), ("tolerance", c_dbl), ("use_fallback_solver", c_int),
                ("reserved", c_int * 5)]


_OP = ctypes.POINTER(_Options)
_matrix_create = {t: _sig(f"sb200_matrix_create_{t}", [c_ptr, c_int, c_int, c_i64, c_i64, c_i64, ctypes.POINTER(c_ptr)])
                  for t in "sdcz"}
_matrix_destroy = _sig("sb200_matrix_destroy", [c_ptr])
_matrix_generate = _sig("sb200_matrix_generate", [c_ptr, c_int, c_i64, c_ptr])
_matrix_from_host = _sig("sb200_matrix_from_host", [c_ptr, c_ptr, c_i64, c_ptr])
_matrix_to_host = _sig("sb200_matrix_to_host", [c_ptr, c_ptr, c_i64, c_ptr])
_matrix_from_host_local = _sig("sb200_matrix_from_host_local", [c_ptr, c_ptr, c_ptr])
_matrix_to_host_local = _sig("sb200_matrix_to_host_local", [c_ptr, c_ptr, c_ptr])
_last_panel_ms = _sig("sb200_last_driver_panel_ms", [c_ptr], c_dbl)
_matrix_copy = _sig("sb200_matrix_copy", [c_ptr, c_ptr, c_ptr])
_matrix_local_tiles = _sig("sb200_matrix_local_tiles", [c_ptr], c_i64)
_matrix_probe_mv = _sig("sb200_matrix_probe_mv", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr])
_matrix_from_scalapack = _sig("sb200_matrix_from_scalapack", [c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr])
_matrix_to_scalapack = _sig("sb200_matrix_to_scalapack", [c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr])
_last_ms = _sig("sb200_last_driver_ms", [c_ptr], c_dbl)
_potrf = {t: _sig(f"sb200_potrf_{t}", [c_ptr, _OP, ctypes.POINTER(c_i64)]) for t in "sdcz"}
_potrf["s_tc05"] = _sig("sb200_potrf_tc05_s", [c_ptr, _OP, ctypes.POINTER(c_i64)])
_potrf_to_host = {t: _sig(f"sb200_potrf_to_host_local_{t}", [c_ptr, _OP, ctypes.POINTER(c_i64), c_ptr]) for t in "sdcz"}
_potrf_stream = {t: _sig(f"sb200_potrf_stream_{t}", [c_ptr, _OP, ctypes.POINTER(c_i64), c_ptr, c_ptr]) for t in "sdcz"}
_gemm = {t: _sig(f"sb200_gemm_{t}", [SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}
_herk = {t: _sig(f"sb200_herk_mat_{t}", [REAL_T[t], c_ptr, REAL_T[t], c_ptr, _OP]) for t in "sdcz"}
_her2k = {t: _sig(f"sb200_her2k_mat_{t}", [SCALAR_T[t], c_ptr, c_ptr, REAL_T[t], c_ptr, _OP]) for t in "sdcz"}
_hemm = {t: _sig(f"sb200_hemm_{t}", [SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}
_potrs = {t: _sig(f"sb200_potrs_{t}", [c_ptr, c_ptr, _OP]) for t in "sdcz"}
_norm_inf = {t: _sig(f"sb200_norm_inf_{t}", [c_ptr, ctypes.POINTER(c_dbl)]) for t in "sdcz"}
_getrf = {"d": _sig("sb200_getrf_d", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)]),
          "s": _sig("sb200_getrf_s", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)]),
          "z": _sig("sb200_getrf_z", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)]),
          "c": _sig("sb200_getrf_c", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)]),
          "s_tc05": _sig("sb200_getrf_tc05_s", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)])}
_getrs = {t: _sig(f"sb200_getrs_{t}", [c_ptr, ctypes.POINTER(c_i64), c_ptr, _OP]) for t in "sdcz"}
_posv_mixed = _sig("sb200_posv_mixed_d", [c_ptr, c_ptr, c_ptr, ctypes.POINTER(_MixedOptions), ctypes.POINTER(c_int),
                                          ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl)])
_gesv_mixed = _sig("sb200_gesv_mixed_d", [c_ptr, ctypes.POINTER(c_i64), c_ptr, c_ptr, ctypes.POINTER(_MixedOptions),
                                          ctypes.POINTER(c_int), ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl)])
_posv_mixed_z = _sig("sb200_posv_mixed_z", [c_ptr, c_ptr, c_ptr, ctypes.POINTER(_MixedOptions), ctypes.POINTER(c_int),
                                            ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl)])
_gesv_mixed_z = _sig("sb200_gesv_mixed_z", [c_ptr, ctypes.POINTER(c_i64), c_ptr, c_ptr, ctypes.POINTER(_MixedOptions),
                                            ctypes.POINTER(c_int), ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl)])
tile_rank = _sig("sb200_tile_rank", [c_int, c_int, c_i64, c_i64])
local_tile_count = _sig("sb200_local_tile_count", [c_int, c_int, c_int, c_int, c_i64, c_i64, c_i64], c_i64)
local_tile_index = _sig("sb200_local_tile_index", [c_int, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_i64], c_i64)


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _opts(opts):
    """sb200_options_t for the caller's dict; keys the caller does not give take the reference's defaults, except
    `lookahead` of potrf, where 0 = the library's tuned depth (see _potrf_opts)."""
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

    def bcast_tiles(self, ranges):
        """Broadcast device ranges in one go (BaseMatrix::listBcast, include/slate/BaseMatrix.hh:1998-2140): `ranges` is
        the same list on every rank of (src_ptr, dst_ptr, nbytes, root) with device pointers (src_ptr is read on the
        root only).  Asynchronous on the current torch stream."""
        n = len(ranges)
        if n == 0:
            return
        src = (c_ptr * n)(*[r[0] for r in ranges]); dst = (c_ptr * n)(*[r[1] for r in ranges])
        nbytes = (ctypes.c_size_t * n)(*[r[2] for r in ranges]); roots = (c_int * n)(*[r[3] for r in ranges])
        check(_bcast_tiles(self._h, n, src, dst, nbytes, roots, _stream()), "bcast_tiles")

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

    def __init__(self, m: int, n: int, nb: int, grid: Grid | None = None, dtype="d"):
        self.grid = grid or default_grid()
        self.m, self.n, self.nb = int(m), int(n), int(nb)
        self.t = dtype_char(dtype)
        self.dtype = np.dtype(NP_DTYPE[self.t])
        h = c_ptr()
        check(_matrix_create[self.t](self.grid._h, ord(self._kind), ord("C"), self.m, self.n, self.nb,
                                     ctypes.byref(h)), "Matrix")
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
        ptr, lda = _host_ptr(hA, self.m, self.n, self.dtype)
        check(_matrix_from_host(self._h, ptr, lda, _stream()), "from_host")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return self

    def to_host(self, out=None):
        import torch
        if out is None:
            out = np.zeros((self.m, self.n), dtype=self.dtype, order="F")
        ptr, lda = _host_ptr(out, self.m, self.n, self.dtype)
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

    def from_scalapack(self, local, lld: int | None = None, sync: bool = True):
        """Gather this rank's tiles from a ScaLAPACK-style local array (Matrix::fromScaLAPACK, include/slate/Matrix.hh:75-99):
        `local` is a torch tensor (CPU or CUDA) holding the column-major local array, leading dimension lld."""
        lld = self._check_scalapack(local, lld)
        check(_matrix_from_scalapack(self._h, local.data_ptr(), lld, int(local.shape[0]), 1 if local.is_cuda else 0, _stream()), "from_scalapack")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return self

    def to_scalapack(self, local, lld: int | None = None, sync: bool = True):
        lld = self._check_scalapack(local, lld)
        check(_matrix_to_scalapack(self._h, local.data_ptr(), lld, int(local.shape[0]), 1 if local.is_cuda else 0, _stream()), "to_scalapack")
        if sync:
            import torch
            torch.cuda.current_stream().synchronize()
        return local

    def _numroc(self):
        """ScaLAPACK numroc: rows and columns of this rank's local array (tile (i, j) at local block (i // p, j // q))."""
        g, nb = self.grid, self.nb
        prow, pcol = g.rank % g.p, g.rank // g.p
        mt, nt = -(-self.m // nb), -(-self.n // nb)
        rows = sum(min(nb, self.m - i * nb) for i in range(prow, mt, g.p))
        cols = sum(min(nb, self.n - j * nb) for j in range(pcol, nt, g.q))
        return rows, cols

    def _check_scalapack(self, t, lld):
        import torch
        if not (isinstance(t, torch.Tensor) and t.dtype == _torch_dtype(self.dtype) and t.is_contiguous() and t.dim() == 2):
            raise Exception_("local array must be a contiguous 2-D torch tensor [local columns][lld] of the matrix type")
        lld = int(lld if lld is not None else t.shape[1])            # row-major [cols][lld] == column-major lld x cols
        rows, cols = self._numroc()
        if lld > t.shape[1]:
            raise Exception_("lld exceeds the local array")
        if lld < max(rows, 1):
            raise Exception_(f"lld = {lld} is smaller than this rank's {rows} local rows")
        if t.shape[0] < cols:
            raise Exception_(f"local array holds {t.shape[0]} columns, this rank owns {cols}")
        return lld

    def _check_local(self, t):
        import torch
        need = self.local_tiles * self.nb * self.nb
        if not (isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.dtype == _torch_dtype(self.dtype)
                and t.is_contiguous() and t.numel() == need):
            raise Exception_(f"local tile buffer must be a contiguous {self.dtype} CPU tensor of {need} elements")

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

    def __init__(self, n: int, nb: int, grid: Grid | None = None, uplo: str = "L", dtype="d"):
        if uplo.upper()[0] != "L":
            raise Exception_("only Uplo::Lower storage is implemented")
        super().__init__(n, n, nb, grid, dtype)


def _torch_dtype(dt):
    import torch
    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[np.dtype(dt)]


def _host_ptr(a, m, n, dtype=np.float64):
    dtype = np.dtype(dtype)
    try:
        import torch
        if isinstance(a, torch.Tensor):
            if a.device.type != "cpu" or a.dtype != _torch_dtype(dtype) or not a.is_contiguous():
                raise Exception_(f"host tensor must be a contiguous {dtype} CPU tensor")
            if tuple(a.shape) != (n, m):
                raise Exception_(f"host tensor holding a column-major {m}x{n} matrix must have shape ({n}, {m})")
            return a.data_ptr(), max(m, 1)
    except ImportError:
        pass
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.shape != (m, n) or not a.flags.f_contiguous:
        raise Exception_(f"host array must be {dtype}, shape ({m}, {n}), Fortran order")
    return a.ctypes.data, max(m, 1)


# -- drivers ---------------------------------------------------------------------------------
def _same_type(*ms):
    t = ms[0].t
    for x in ms[1:]:
        if x.t != t:
            raise Exception_("operands must have the same element type")
    return t


_gemm_op = {t: _sig(f"sb200_gemm_op_{t}", [c_int, c_int, SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


def gemm(alpha, A: Matrix, B: Matrix, beta, C: Matrix, opts: dict | None = None, opA: str = "N", opB: str = "N"):
    """C = alpha op(A) op(B) + beta C  (slate::gemm, src/gemm.cc:82-105 -> gemmC).  opA / opB "T" | "C" stand for the
    (conjugate-)transposed views slate::gemm is handed: A is then the stored k x m matrix, B the stored n x k one."""
    t = _same_type(A, B, C)
    o = _opts(opts)
    if opA == "N" and opB == "N":
        check(_gemm[t](scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "gemm")
    else:
        check(_gemm_op[t](ord(opA), ord(opB), scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "gemm")


multiply = gemm     # simplified API name (include/slate/simplified_api.hh)


_herk_op = {t: _sig(f"sb200_herk_op_{t}", [c_int, REAL_T[t], c_ptr, REAL_T[t], c_ptr, _OP]) for t in "sdcz"}
_her2k_op = {t: _sig(f"sb200_her2k_op_{t}", [c_int, SCALAR_T[t], c_ptr, c_ptr, REAL_T[t], c_ptr, _OP]) for t in "sdcz"}
_syrk_op = {t: _sig(f"sb200_syrk_op_{t}", [c_int, SCALAR_T[t], c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}
_syr2k_op = {t: _sig(f"sb200_syr2k_op_{t}", [c_int, SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


def herk(alpha: float, A: Matrix, beta: float, C: "HermitianMatrix", opts: dict | None = None, op: str = "N"):
    """C = alpha A A^H + beta C, C Hermitian (lower), alpha / beta real (slate::herk, src/herk.cc:25-162;
    for real types this is syrk, as in the reference).  op "C": A is the stored k x n matrix whose conjugate-transposed
    view slate::herk is handed, C = alpha A^H A + beta C."""
    t = _same_type(A, C)
    o = _opts(opts)
    if op == "N":
        check(_herk[t](REAL_T[t](float(alpha)), A._h, REAL_T[t](float(beta)), C._h, ctypes.byref(o)), "herk")
    else:
        check(_herk_op[t](ord(op), REAL_T[t](float(alpha)), A._h, REAL_T[t](float(beta)), C._h, ctypes.byref(o)), "herk")


rank_k_update = herk


def her2k(alpha, A: Matrix, B: Matrix, beta: float, C: "HermitianMatrix", opts: dict | None = None, op: str = "N"):
    """C = alpha A B^H + conj(alpha) B A^H + beta C, C Hermitian (lower), beta real (slate::her2k, src/her2k.cc:27-170;
    for real types this is syr2k).  op "C": A, B stored k x n, C = alpha A^H B + conj(alpha) B^H A + beta C."""
    t = _same_type(A, B, C)
    o = _opts(opts)
    if op == "N":
        check(_her2k[t](scalar(t, alpha), A._h, B._h, REAL_T[t](float(beta)), C._h, ctypes.byref(o)), "her2k")
    else:
        check(_her2k_op[t](ord(op), scalar(t, alpha), A._h, B._h, REAL_T[t](float(beta)), C._h, ctypes.byref(o)), "her2k")


rank_2k_update = her2k

_syrk = {t: _sig(f"sb200_syrk_mat_{t}", [SCALAR_T[t], c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}
_syr2k = {t: _sig(f"sb200_syr2k_mat_{t}", [SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


def syrk(alpha, A: Matrix, beta, C: "HermitianMatrix", opts: dict | None = None, op: str = "N"):
    """C = alpha A A^T + beta C, C symmetric (lower tiles), no conjugation (slate::syrk, src/syrk.cc).
    op "T": A stored k x n (its transposed view is what slate::syrk is handed), C = alpha A^T A + beta C."""
    t = _same_type(A, C)
    o = _opts(opts)
    if op == "N":
        check(_syrk[t](scalar(t, alpha), A._h, scalar(t, beta), C._h, ctypes.byref(o)), "syrk")
    else:
        check(_syrk_op[t](ord(op), scalar(t, alpha), A._h, scalar(t, beta), C._h, ctypes.byref(o)), "syrk")


def syr2k(alpha, A: Matrix, B: Matrix, beta, C: "HermitianMatrix", opts: dict | None = None, op: str = "N"):
    """C = alpha A B^T + alpha B A^T + beta C, C symmetric (lower tiles), no conjugation (slate::syr2k, src/syr2k.cc).
    op "T": A, B stored k x n, C = alpha A^T B + alpha B^T A + beta C."""
    t = _same_type(A, B, C)
    o = _opts(opts)
    if op == "N":
        check(_syr2k[t](scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "syr2k")
    else:
        check(_syr2k_op[t](ord(op), scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "syr2k")


_hemm_side = {t: _sig(f"sb200_hemm_side_{t}", [c_int, SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


def hemm(alpha, A: "HermitianMatrix", B: Matrix, beta, C: Matrix, opts: dict | None = None, side: str = "L"):
    """C = alpha A B + beta C (side "L") or C = alpha B A + beta C (side "R") with A Hermitian, lower tiles
    (slate::hemm(side, ...), src/hemmC.cc)."""
    t = _same_type(A, B, C)
    o = _opts(opts)
    if side == "L":
        check(_hemm[t](scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "hemm")
    else:
        check(_hemm_side[t](ord(side), scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "hemm")


_symm = {t: _sig(f"sb200_symm_{t}", [SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


_symm_side = {t: _sig(f"sb200_symm_side_{t}", [c_int, SCALAR_T[t], c_ptr, c_ptr, SCALAR_T[t], c_ptr, _OP]) for t in "sdcz"}


def symm(alpha, A: "HermitianMatrix", B: Matrix, beta, C: Matrix, opts: dict | None = None, side: str = "L"):
    """C = alpha A B + beta C (side "L") or C = alpha B A + beta C (side "R") with A (complex-)symmetric (lower tiles), no
    conjugation (slate::symm(side, ...), src/symm.cc)."""
    t = _same_type(A, B, C)
    o = _opts(opts)
    if side == "L":
        check(_symm[t](scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "symm")
    else:
        check(_symm_side[t](ord(side), scalar(t, alpha), A._h, B._h, scalar(t, beta), C._h, ctypes.byref(o)), "symm")


_trmm = {t: _sig(f"sb200_trmm_{t}", [c_int, c_int, c_int, c_int, SCALAR_T[t], c_ptr, c_ptr, _OP]) for t in "sdcz"}


def trmm(alpha, A: "HermitianMatrix", B: Matrix, side: str = "L", uplo: str = "L", op: str = "N", diag: str = "N",
         opts: dict | None = None):
    """B = alpha op(A) B (side "L") or B = alpha B op(A) (side "R") with A lower triangular (the lower tiles of A) and
    op "N" | "T" | "C" the transposed view slate::trmm would be handed (slate::trmm(side, alpha, A, B), src/trmm.cc)."""
    t = _same_type(A, B)
    o = _opts(opts)
    check(_trmm[t](ord(side), ord(uplo), ord(op), ord(diag), scalar(t, alpha), A._h, B._h, ctypes.byref(o)), "trmm")


triangular_multiply = trmm


_norm = {t: _sig(f"sb200_norm_{t}", [c_int, c_int, c_ptr, ctypes.POINTER(c_dbl)]) for t in "sdcz"}
_NORM_CHAR = {"max": "M", "one": "O", "1": "O", "inf": "I", "fro": "F", "M": "M", "O": "O", "I": "I", "F": "F"}


def norm(kind: str, A: Matrix, symmetric: bool = False) -> float:
    """slate::norm(Norm::Max | One | Inf | Fro, A) (src/norm.cc): kind "max" | "one" | "inf" | "fro"; A a general Matrix, or a
    HermitianMatrix handle whose lower tiles are read as a Hermitian matrix (default) or, with symmetric=True, as a
    (complex-)symmetric one (slate::SymmetricMatrix)."""
    if kind not in _NORM_CHAR:
        raise Exception_(f"unknown norm {kind!r}")
    v = c_dbl(0.0)
    flavour = ("S" if symmetric else "H") if A._kind == "H" else "G"
    check(_norm[A.t](ord(_NORM_CHAR[kind]), ord(flavour), A._h, ctypes.byref(v)), "norm")
    return float(v.value)


def norm_inf(A: Matrix) -> float:
    """slate::norm(Norm::Inf, A) for a general or Hermitian matrix."""
    v = c_dbl(0.0)
    check(_norm_inf[A.t](A._h, ctypes.byref(v)), "norm")
    return float(v.value)


def potrf(A: HermitianMatrix, opts: dict | None = None, out_local=None, in_local=None) -> int:
    """Cholesky A = L L^H, lower (slate::potrf, src/potrf.cc:262-281).
    out_local (optional): packed host tile buffer as for Matrix.to_host_local; every finished block column is copied
    into it while the factorisation runs (pinned memory: the D2H overlaps the trailing updates).
    in_local (optional, one rank): packed host tile buffer as for Matrix.from_host_local; the matrix streams in by
    chunks of block columns while earlier chunks are factored (bitwise the same factor).
    Returns info: 0, or i > 0 if the leading minor of order i is not positive definite.
    opts['tensor_core_fp32'] (float matrices only): run the trailing update on the tcgen05
    FP32-emulated kernel, as posv_mixed does for its low-precision factorisation."""
    o = _opts(opts)
    if "lookahead" not in (opts or {}):
        o.lookahead = 0                     # the library's tuned depth (csrc/runtime_internal.hh POTRF_DEFAULT_LOOKAHEAD)
    info = c_i64(0)
    key = A.t
    if (opts or {}).get("tensor_core_fp32"):
        if A.t != "s":
            raise Exception_("tensor_core_fp32 applies to float matrices")
        key = "s_tc05"
    if in_local is not None:
        if key == "s_tc05":
            raise Exception_("in_local is not combined with tensor_core_fp32")
        A._check_local(in_local)
        if out_local is not None:
            A._check_local(out_local)
        check(_potrf_stream[A.t](A._h, ctypes.byref(o), ctypes.byref(info), in_local.data_ptr(),
                                 out_local.data_ptr() if out_local is not None else None), "potrf")
        return int(info.value)
    if out_local is not None:
        if key == "s_tc05":
            raise Exception_("out_local is not combined with tensor_core_fp32")
        A._check_local(out_local)
        check(_potrf_to_host[A.t](A._h, ctypes.byref(o), ctypes.byref(info), out_local.data_ptr()), "potrf")
        return int(info.value)
    check(_potrf[key](A._h, ctypes.byref(o), ctypes.byref(info)), "potrf")
    return int(info.value)


chol_factor = potrf


def potrs(A: HermitianMatrix, B: Matrix, opts: dict | None = None):
    """Solve A X = B with the Cholesky factor from potrf; B is overwritten by X (slate::potrs, src/potrs.cc)."""
    t = _same_type(A, B)
    o = _opts(opts)
    check(_potrs[t](A._h, B._h, ctypes.byref(o)), "potrs")


chol_solve_using_factor = potrs


def _pivots_flat(pivots):
    flat = [v for blk in pivots for pr in blk for v in pr]
    return (c_i64 * max(len(flat), 1))(*flat)


def getrf(A: Matrix, opts: dict | None = None):
    """LU with partial pivoting P A = L U (slate::getrf, src/getrf.cc:320-350).
    Returns (pivots, info); pivots[k] = list of (tileIndex, elementOffset) relative to the
    panel sub-matrix A(k:mt-1, k), as slate::Pivots (include/slate/types.hh:84-117)."""
    o = _opts(opts)
    info = c_i64(0)
    mn = min(A.m, A.n)
    flat = (c_i64 * (2 * max(mn, 1)))()
    key = A.t
    if (opts or {}).get("tensor_core_fp32"):
        if A.t != "s":
            raise Exception_("tensor_core_fp32 applies to float matrices")
        key = "s_tc05"
    if key not in _getrf:
        raise Exception_(f"getrf is not implemented for {A.dtype}")
    check(_getrf[key](A._h, flat, ctypes.byref(o), ctypes.byref(info)), "getrf")
    piv = np.frombuffer(flat, dtype=np.int64)[: 2 * mn].reshape(-1, 2)
    nb = A.nb
    pivots = [list(map(tuple, piv[k0:min(k0 + nb, mn)].tolist())) for k0 in range(0, mn, nb)]      # (tileIndex, elementOffset) ints
    return pivots, int(info.value)


_getrf_tntpiv = {t: _sig(f"sb200_getrf_tntpiv_{t}", [c_ptr, ctypes.POINTER(c_i64), _OP, ctypes.POINTER(c_i64)]) for t in "sdcz"}


def getrf_tntpiv(A: Matrix, opts: dict | None = None):
    """LU with tournament pivoting, CALU (slate::getrf_tntpiv, src/getrf_tntpiv.cc).  Returns (pivots, info) as getrf.
    The participants of a panel's tournament are the process rows of A's grid."""
    if A.t not in _getrf_tntpiv:
        raise Exception_(f"getrf_tntpiv is implemented for float and double, not {A.dtype}")
    o = _opts(opts)
    info = c_i64(0)
    mn = min(A.m, A.n)
    flat = (c_i64 * (2 * max(mn, 1)))()
    check(_getrf_tntpiv[A.t](A._h, flat, ctypes.byref(o), ctypes.byref(info)), "getrf_tntpiv")
    piv = np.frombuffer(flat, dtype=np.int64)[: 2 * mn].reshape(-1, 2)
    nb = A.nb
    pivots = [list(map(tuple, piv[k0:min(k0 + nb, mn)].tolist())) for k0 in range(0, mn, nb)]
    return pivots, int(info.value)


def lu_factor(A: Matrix, opts: dict | None = None):
    """slate::lu_factor (src/getrf.cc:320-350): Option::MethodLU picks partial pivoting (default), CALU or NoPiv;
    NoPiv returns an empty pivot list, as the reference leaves `pivots` untouched."""
    method = str((opts or {}).get("method_lu", "PPLU")).lower()
    if method in ("pplu", "partialpiv", "auto"):
        return getrf(A, opts)
    if method == "calu":
        return getrf_tntpiv(A, opts)
    if method == "nopiv":
        return [], getrf_nopiv(A, opts)
    raise Exception_(f"unknown value for MethodLU: {method}")


_getrf_nopiv = {t: _sig(f"sb200_getrf_nopiv_{t}", [c_ptr, _OP, ctypes.POINTER(c_i64)]) for t in "sdcz"}


def getrf_nopiv(A: Matrix, opts: dict | None = None) -> int:
    """LU without pivoting A = L U (slate::getrf_nopiv, src/getrf_nopiv.cc).  Returns info (first zero pivot + 1, or 0)."""
    if A.t not in _getrf_nopiv:
        raise Exception_(f"getrf_nopiv is implemented for float and double, not {A.dtype}")
    o = _opts(opts)
    info = c_i64(0)
    check(_getrf_nopiv[A.t](A._h, ctypes.byref(o), ctypes.byref(info)), "getrf_nopiv")
    return int(info.value)


lu_factor_nopiv = getrf_nopiv


_getrs_op = _sig("sb200_getrs_op", [c_int, c_ptr, ctypes.POINTER(c_i64), c_ptr, _OP])


def getrs(A: Matrix, pivots, B: Matrix, opts: dict | None = None, op: str = "N"):
    """Solve A X = B with the LU factors and pivots from getrf; B is overwritten (slate::getrs, src/getrs.cc).
    op "T" | "C": solve op(A) X = B, as slate::getrs does when it is handed a (conjugate-)transposed view (:97-112)."""
    if op != "N":
        _same_type(A, B)
        o = _opts(opts)
        check(_getrs_op(ord(op), A._h, _pivots_flat(pivots), B._h, ctypes.byref(o)), "getrs")
        return
    t = _same_type(A, B)
    if t not in _getrs:
        raise Exception_(f"getrs is not implemented for {A.dtype}")
    o = _opts(opts)
    check(_getrs[t](A._h, _pivots_flat(pivots), B._h, ctypes.byref(o)), "getrs")


lu_solve_using_factor = getrs


def posv(A: HermitianMatrix, B: Matrix, opts: dict | None = None) -> int:
    """Solve A X = B, A Hermitian positive definite: potrf, then potrs if the factorisation succeeded; A holds the
    factor, B the solution (slate::posv / chol_solve, src/posv.cc:80-94).  Returns info."""
    info = potrf(A, opts)
    if info == 0:
        potrs(A, B, opts)
    return info


def gesv(A: Matrix, B: Matrix, opts: dict | None = None):
    """Solve A X = B: getrf, then getrs if no pivot was exactly zero; A holds L and U, B the solution
    (slate::gesv / lu_solve, src/gesv.cc:95-109).  Returns (pivots, info)."""
    pivots, info = getrf(A, opts)
    if info == 0:
        getrs(A, pivots, B, opts)
    return pivots, info


chol_solve = posv
lu_solve = gesv

_trsm_mat = {t: _sig(f"sb200_trsm_mat_{t}", [c_int, c_int, c_int, c_int, SCALAR_T[t], c_ptr, c_ptr, _OP]) for t in "sdcz"}


def trsm(alpha, A: Matrix, B: Matrix, side: str = "L", uplo: str = "L", op: str = "N", diag: str = "N",
         opts: dict | None = None):
    """B = alpha op(A)^-1 B (side "L") or B = alpha B op(A)^-1 (side "R") with A triangular: the lower tiles of a
    HermitianMatrix handle, or the lower / upper triangle of a square Matrix (an LU factor); op "N" | "T" | "C" is the
    transposed view slate::trsm would be handed (slate::trsm(side, alpha, A, B) / triangular_solve, src/trsm.cc)."""
    t = _same_type(A, B)
    o = _opts(opts)
    check(_trsm_mat[t](ord(side), ord(uplo), ord(op), ord(diag), scalar(t, alpha), A._h, B._h, ctypes.byref(o)), "trsm")


triangular_solve = trsm


def getrs_nopiv(A: Matrix, B: Matrix, opts: dict | None = None):
    """Solve A X = B with the factors of getrf_nopiv: forward substitution with the unit lower triangle of A, backward
    substitution with its upper triangle -- the two sweeps getrs runs, without the row permutation
    (slate::getrs_nopiv, src/getrs_nopiv.cc:20-53)."""
    trsm(1.0, A, B, side="L", uplo="L", op="N", diag="U", opts=opts)
    trsm(1.0, A, B, side="L", uplo="U", op="N", diag="N", opts=opts)


def gesv_nopiv(A: Matrix, B: Matrix, opts: dict | None = None) -> int:
    """getrf_nopiv, then getrs_nopiv when no pivot was zero (slate::gesv_nopiv, src/gesv_nopiv.cc).  Returns info."""
    info = getrf_nopiv(A, opts)
    if info == 0:
        getrs_nopiv(A, B, opts)
    return info


lu_solve_using_factor_nopiv = getrs_nopiv
lu_solve_nopiv = gesv_nopiv

MIXED_TIMERS = ("total", "factor_lo", "solve_lo", "residual_hi", "add_hi", "factor_hi", "solve_hi", "norm_convert")


def _mixed_opts(opts):
    mo = _MixedOptions()
    mo.max_iterations = int((opts or {}).get("max_iterations", 30))
    mo.tolerance = float((opts or {}).get("tolerance", 0.0))
    mo.use_fallback_solver = int(bool((opts or {}).get("use_fallback_solver", True)))
    return mo


def posv_mixed(A: HermitianMatrix, B: Matrix, X: Matrix, opts: dict | None = None):
    """Mixed-precision Cholesky solve (slate::posv_mixed<double,float>, src/posv_mixed.cc:111-297).
    Returns (info, iter, timers_ms): iter as the reference (>= 0 refinement steps, -3 low-precision
    factor failed, -(max_iterations+1) not converged -> FP64 fallback, which overwrites A)."""
    mo = _mixed_opts(opts)
    it, info, tm = c_int(0), c_i64(0), (c_dbl * 8)()
    if A.t not in "dz":
        raise Exception_(f"posv_mixed takes double or complex<double> matrices, not {A.dtype}")
    f = _posv_mixed_z if A.t == "z" else _posv_mixed          # <complex<double>, complex<float>> | <double, float>
    check(f(A._h, B._h, X._h, ctypes.byref(mo), ctypes.byref(it), ctypes.byref(info), tm), "posv_mixed")
    return int(info.value), int(it.value), dict(zip(MIXED_TIMERS, list(tm)))


def gesv_mixed(A: Matrix, B: Matrix, X: Matrix, opts: dict | None = None):
    """Mixed-precision LU solve (slate::gesv_mixed<double,float>, src/gesv_mixed.cc:106-300).
    Returns (info, iter, pivots, timers_ms)."""
    mo = _mixed_opts(opts)
    it, info, tm = c_int(0), c_i64(0), (c_dbl * 8)()
    mn = min(A.m, A.n)
    flat = (c_i64 * (2 * max(mn, 1)))()
    if A.t not in "dz":
        raise Exception_(f"gesv_mixed takes double or complex<double> matrices, not {A.dtype}")
    f = _gesv_mixed_z if A.t == "z" else _gesv_mixed
    check(f(A._h, flat, B._h, X._h, ctypes.byref(mo), ctypes.byref(it), ctypes.byref(info), tm), "gesv_mixed")
    piv = np.frombuffer(flat, dtype=np.int64)[: 2 * mn].reshape(-1, 2)
    pivots = [[(int(t), int(off)) for t, off in piv[k0:min(k0 + A.nb, mn)]] for k0 in range(0, mn, A.nb)]
    return int(info.value), int(it.value), pivots, dict(zip(MIXED_TIMERS, list(tm)))


# -- probe-vector residual checks (bench-size substitute for the tester's ||B - A X|| checks) -------------
def probe_mv(A: Matrix, x, part: str = "G", op: str = "N", diag: str = "N", use_abs: bool = False):
    """y = op(part(A)) x for a replicated device vector x (torch CUDA tensor of the matrix type): every rank multiplies
    the tiles it stores (sb200_matrix_probe_mv) and the partial results are summed over the grid's ranks."""
    import torch
    rows, cols = (A.m, A.n) if op == "N" else (A.n, A.m)
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == _torch_dtype(A.dtype) and x.is_contiguous()
            and x.numel() == cols):
        raise Exception_(f"probe vector must be a contiguous CUDA {A.dtype} tensor of {cols} elements")
    y = torch.zeros(rows, dtype=x.dtype, device=x.device)
    torch.cuda.current_stream().synchronize()
    check(_matrix_probe_mv(A._h, ord(part), ord(op), ord(diag), 1 if use_abs else 0, x.data_ptr(), y.data_ptr(), _stream()),
          "probe_mv")
    if A.grid.p * A.grid.q > 1:
        import torch.distributed as dist
        if y.is_complex():
            yr = torch.view_as_real(y)
            dist.all_reduce(yr)
        else:
            dist.all_reduce(y)
    return y


def _probe_vector(n, dtype, seed=7):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, dtype=torch.float64, generator=g) - 0.5
    if np.dtype(dtype).kind == "c":
        x = torch.complex(x, torch.rand(n, dtype=torch.float64, generator=g) - 0.5)
    return x.to(_torch_dtype(dtype)).cuda()


def _eps(dtype):
    return float(np.finfo(np.dtype(dtype)).eps)


def potrf_residual(A0: HermitianMatrix, L: HermitianMatrix, seed: int = 7) -> dict:
    """|| A x - L (L^H x) ||_max / (n ||A||_1 ||x||_max) for a seeded probe vector (the tester's Cholesky check,
    test/test_posv.cc:336-342, with X = one vector); tolerance = the tester's 50 eps / 2."""
    x = _probe_vector(A0.n, A0.dtype, seed)
    y = probe_mv(A0, x, "H")
    z = probe_mv(L, x, "L", "C")
    w = probe_mv(L, z, "L", "N")
    anorm = float(probe_mv(A0, x.new_ones(A0.n), "H", use_abs=True).abs().max())
    err = float((y - w).abs().max()) / (max(A0.n, 1) * anorm * float(x.abs().max()))
    tol = 25.0 * _eps(A0.dtype) if np.dtype(A0.dtype).kind != "c" else 25.0 * _eps(np.dtype(A0.dtype).char.lower())
    return {"residual": err, "tol": tol, "pass": bool(err <= tol), "kind": "||A x - L (L^H x)|| / (n ||A||_1 ||x||)"}


def getrf_residual(A0: Matrix, LU: Matrix, pivots, seed: int = 7) -> dict:
    """|| P A x - L (U x) ||_max / (n ||A||_1 ||x||_max) for a seeded probe vector (the tester's LU check,
    test/test_gesv.cc:371-377, with X = one vector); pivots as returned by getrf."""
    import torch
    x = _probe_vector(A0.n, A0.dtype, seed)
    y = probe_mv(A0, x, "G").cpu().numpy()
    nb = A0.nb
    for k, blk in enumerate(pivots):                      # apply P: the panels' interchanges in order
        for j, (t, off) in enumerate(blk):
            r, p = k * nb + j, (k + t) * nb + off
            if p != r:
                y[r], y[p] = y[p], y[r]
    z = probe_mv(LU, x, "U")                              # U is min(m,n) x n: rows beyond it come out zero
    mn = min(LU.m, LU.n)
    zz = torch.zeros(LU.n, dtype=z.dtype, device=z.device)
    zz[:mn] = z[:mn]
    w = probe_mv(LU, zz, "L", "N", "U").cpu().numpy()
    anorm = float(probe_mv(A0, x.new_ones(A0.m), "G", "C", use_abs=True).abs().max())      # column sums: one-norm
    err = float(np.abs(y - w).max()) / (max(A0.n, 1) * anorm * float(x.abs().max()))
    tol = 25.0 * _eps(np.dtype(A0.dtype).char.lower() if np.dtype(A0.dtype).kind == "c" else A0.dtype)
    return {"residual": err, "tol": tol, "pass": bool(err <= tol), "kind": "||P A x - L (U x)|| / (n ||A||_1 ||x||)"}


def gemm_residual(alpha, A: Matrix, B: Matrix, beta, c0x, C: Matrix, x) -> dict:
    """|| C x - (alpha A (B x) + beta C0 x) ||_max / (sqrt(k) (|alpha| ||A|| ||B|| + |beta| ||C0||) ||x||)-style check with a
    probe vector (test/test_gemm.cc:205-207 with X = one vector).  c0x = probe_mv(C0, x) taken before the multiply."""
    bx = probe_mv(B, x, "G")
    abx = probe_mv(A, bx, "G")
    cx = probe_mv(C, x, "G")
    want = alpha * abx + beta * c0x
    ones_n, ones_k = x.new_ones(B.n), x.new_ones(A.n)
    an = float(probe_mv(A, ones_k, "G", use_abs=True).abs().max())
    bn = float(probe_mv(B, ones_n, "G", use_abs=True).abs().max())
    c0n = float(c0x.abs().max())
    denom = (abs(alpha) * an * bn + abs(beta) * c0n) * float(x.abs().max()) * np.sqrt(float(A.n))
    err = float((cx - want).abs().max()) / denom
    tol = 3.0 * _eps(np.dtype(A.dtype).char.lower() if np.dtype(A.dtype).kind == "c" else A.dtype)
    return {"residual": err, "tol": tol, "pass": bool(err <= tol), "kind": "||C x - (alpha A (B x) + beta C0 x)|| / (sqrt(k) (|alpha| ||A|| ||B|| + |beta| ||C0 x||) ||x||)"}
