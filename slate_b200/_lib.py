"""ctypes binding of libslate_b200.so (the C ABI declared in include/slate_b200.h).

There is NO CPU fallback: if the shared library is missing the import fails loudly,
and every compute entry point returns SB200_ENODEV / a CUDA error without a GPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libslate_b200.so")


class SB200Error(RuntimeError):
    pass


def _load():
    # libslate_b200.so needs libnccl.so.2.  torch bundles a NEWER NCCL than the system one and
    # refuses to import once the older library is resident (undefined ncclDevCommCreate), so make
    # sure torch's copy is the one the process loads: import torch first, when it is installed.
    try:
        import torch  # noqa: F401
    except ImportError:
        pass
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C slate_b200/csrc`. slate_b200 has no CPU/PyTorch fallback.")
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_flt = ctypes.c_float
c_ptr = ctypes.c_void_p


class c64(ctypes.Structure):
    _fields_ = [("re", c_dbl), ("im", c_dbl)]


class c32(ctypes.Structure):
    _fields_ = [("re", c_flt), ("im", c_flt)]


lib.sb200_strerror.restype = ctypes.c_char_p
lib.sb200_strerror.argtypes = [c_int]
lib.sb200_launch_count.restype = c_i64
lib.sb200_version.restype = c_int


def check(code, what=""):
    if code != 0:
        msg = lib.sb200_strerror(int(code)).decode()
        raise SB200Error(f"{what or 'slate_b200 call'} failed: {msg} (code {code})")


def _sig(name, argtypes, restype=c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


def scalar(dtype_char, v):
    """Pack a Python scalar as the C ABI scalar of the given type suffix."""
    if dtype_char == "d":
        return c_dbl(float(v.real if isinstance(v, complex) else v))
    if dtype_char == "s":
        return c_flt(float(v.real if isinstance(v, complex) else v))
    v = complex(v)
    return c64(v.real, v.imag) if dtype_char == "z" else c32(v.real, v.imag)


SCALAR_T = {"s": c_flt, "d": c_dbl, "c": c32, "z": c64}
REAL_T = {"s": c_flt, "d": c_dbl, "c": c_flt, "z": c_dbl}
