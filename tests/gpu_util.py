"""Helpers for the -m gpu parity tests: torch is used only to hold device memory and streams."""
import ctypes

import numpy as np

from slate_b200._lib import lib, check, c_i64, c_int, c_dbl, c_flt, c_ptr, c64, c32

NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
REAL = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def sync():
    import torch
    torch.cuda.synchronize()


class DevTiles:
    """A batch of column-major tiles on the device (each its own allocation, so that pointer
    arrays are genuinely scattered) + the device pointer array."""

    def __init__(self, arrays):
        import torch
        self.shapes = [a.shape for a in arrays]
        self.dtype = arrays[0].dtype
        self.t = []
        for a in arrays:
            f = np.asfortranarray(a)
            buf = torch.from_numpy(np.frombuffer(f.tobytes(order="F"), dtype=np.uint8).copy()).cuda()
            self.t.append(buf)
        self.ptrs = torch.tensor([b.data_ptr() for b in self.t], dtype=torch.int64, device="cuda")

    @property
    def p(self):
        return self.ptrs.data_ptr()

    def get(self):
        sync()
        out = []
        for b, shp in zip(self.t, self.shapes):
            raw = b.cpu().numpy().tobytes()
            out.append(np.frombuffer(raw, dtype=self.dtype).reshape(shp, order="F").copy())
        return out


def dev_zeros(n, dtype):
    import torch
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}[dtype]
    return torch.zeros(n, dtype=tdt, device="cuda")


def scal(t, v):
    if t == "d":
        return c_dbl(float(np.real(v)))
    if t == "s":
        return c_flt(float(np.real(v)))
    v = complex(v)
    return c64(v.real, v.imag) if t == "z" else c32(v.real, v.imag)


SC = {"s": c_flt, "d": c_dbl, "c": c32, "z": c64}
RSC = {"s": c_flt, "d": c_dbl, "c": c_flt, "z": c_dbl}


def fn(name, argtypes):
    f = getattr(lib, name)
    f.argtypes = argtypes
    f.restype = c_int
    return f


def rng_tiles(rng, batch, m, n, t):
    dt = NP[t]
    out = []
    for _ in range(batch):
        a = rng.random((m, n))
        if t in "cz":
            a = a + 1j * rng.random((m, n))
        out.append(np.asfortranarray(a.astype(dt)))
    return out


# LU factor against the oracle / the reference's golden factor, relative to the largest entry, for n <= 2048 with
# IDENTICAL pivots: observed 1e-13 .. 4e-13 on B200 (1-, 2-, 4- and 8-GPU grids, every base-kernel variant); the bound
# leaves a factor of 5.  (Round 1 used 1e-11.)
GETRF_TOL = 2e-12
