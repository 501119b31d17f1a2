"""Run the REAL host code of libslate_b200.so -- the drivers behind the pending GPU tests -- on the CPU, without a GPU:

    g++ -O1 -std=c++17 -shared -fPIC -I/usr/local/cuda/include scratch/cpu_standin/fake_cudart.cc -o /tmp/libfakecudart.so
    LD_PRELOAD=/tmp/libfakecudart.so python scratch/cpu_standin/run_real_host_code_on_cpu.py [-k substring]

fake_cudart.cc stands in for libcudart: device memory is host memory and every kernel the drivers launch (tile GEMMs, the
diagonal-tile fill kernels, the 64 x 64 triangular inverses, scale, row gather) is replaced by a plain loop that computes
what the kernel is specified to compute.  The test functions of tests/test_zzzzz_gpu_blas3_variants.py then run against the
real `slate_b200.host` -> C ABI -> solve.cu: plan building, tile indices, operand roles, launch order, in-place updates are
the shipped code; only the arithmetic inside a launch is emulated.  SB200_TRSM_FUSED=0 / SB200_TRSM_SMALL=0 select the
launch chains of the triangular solves that consist of those emulated kernels (the one-launch forms are validated on B200s).
Matrix.generate is replaced by from_host of the oracle's generator (the Philox kernel is not emulated).
Test infrastructure: nothing in the product or the test suite imports this."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SB200_TRSM_FUSED", "0")
os.environ.setdefault("SB200_TRSM_SMALL", "0")

if "libfakecudart" not in os.environ.get("LD_PRELOAD", ""):
    sys.exit("run with LD_PRELOAD=/tmp/libfakecudart.so (see the docstring)")

# a torch that is only a stream token: host.py asks it for the current stream and synchronises it
torch = types.ModuleType("torch")


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


class _Tensor:            # isinstance(x, torch.Tensor) is False for numpy arrays
    pass


torch.Tensor = _Tensor
torch.cuda = types.SimpleNamespace(current_stream=lambda: _Stream(), set_device=lambda d: None, is_available=lambda: True,
                                   current_device=lambda: 0)
sys.modules["torch"] = torch

import numpy as np                                           # noqa: E402
from oracle import slate_oracle as o                         # noqa: E402
import slate_b200.host as sl                                 # noqa: E402


def _generate(self, kind, seed):
    return self.from_host(np.asfortranarray(o.generate(kind, self.m, self.n, seed, self.dtype)))


sl.Matrix.generate = _generate


# The factorisations themselves are validated on B200s and their kernels are not emulated: the routines under test here
# (getrs with an op, the posv / gesv / gesv_nopiv wrappers) get their factors from the oracle, uploaded through from_host.
def _wide(a):
    return a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)


def _getrf(A, opts=None):
    LU, piv, info = o.getrf(_wide(A.to_host()), A.nb)
    A.from_host(np.asfortranarray(LU.astype(A.dtype)))
    return piv, info


def _getrf_nopiv(A, opts=None):
    LU, info = o.getrf_nopiv(_wide(A.to_host()), A.nb)
    A.from_host(np.asfortranarray(LU.astype(A.dtype)))
    return info


def _potrf(A, opts=None, out_local=None, in_local=None):
    L, info = o.potrf(o.he_full(np.tril(_wide(A.to_host()))), A.nb)
    if info == 0:
        A.from_host(np.asfortranarray(np.tril(L).astype(A.dtype)))
    return info


sl.getrf, sl.getrf_nopiv, sl.potrf = _getrf, _getrf_nopiv, _potrf
os.environ.setdefault("SB200_GEMM_BT", "0")                  # plain 'N','N' gemm without the (unemulated) transpose kernel
import ctypes                                                # noqa: E402
_rt = ctypes.CDLL(None)

import check_gpu_test_logic as runner                        # noqa: E402


def _without_norm_inf(f):
    def g(**kw):
        saved = sl.norm_inf
        sl.norm_inf = lambda A: sl.norm("inf", A)
        try:
            return f(**kw)
        finally:
            sl.norm_inf = saved
    g.__code__ = f.__code__ if False else g.__code__
    g.pytestmark = getattr(f, "pytestmark", [])
    g.argnames = f.__code__.co_varnames[:f.__code__.co_argcount]
    return g


def main():
    import importlib
    import itertools
    mod = importlib.import_module("tests.test_zzzzz_gpu_blas3_variants")
    golden = os.path.join(ROOT, "tests", "golden")
    sel = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "-k" else ""
    ran = failed = 0
    for name in sorted(n for n in dir(mod) if n.startswith("test_")):
        f = getattr(mod, name)
        if (sel and sel not in name) or "reference_tester" in name:
            continue
        if name == "test_general_norms_vs_numpy":
            f = _without_norm_inf(f)                         # sl.norm_inf's one-kernel row sums (validated on B200s) are not emulated
        marks = [m for m in getattr(f, "pytestmark", []) if m.name == "parametrize"]
        axes = []
        for m in marks:
            names = [x.strip() for x in m.args[0].split(",")]
            axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in m.args[1]])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo:
                kw.update(d)
            if max(kw.get("n", 0), kw.get("m", 0), kw.get("k", 0)) > int(os.environ.get("STANDIN_MAX_DIM", "520")) \
                    and not os.environ.get("STANDIN_ALL"):
                continue                                     # the emulated GEMM is a triple loop: keep to the small and ragged shapes
            argnames = getattr(f, "argnames", None) or f.__code__.co_varnames[:f.__code__.co_argcount]
            if "sl" in argnames:
                kw["sl"] = sl
            if "golden_dir" in argnames:
                kw["golden_dir"] = golden
            ran += 1
            try:
                f(**kw)
            except Exception as ex:   # noqa: BLE001
                _rt.cudaGetLastError()                       # an unconsumed launch error must not leak into the next case
                failed += 1
                print(f"FAIL {name} {({k: v for k, v in kw.items() if k not in ('sl', 'golden_dir')})}: {type(ex).__name__}: {str(ex)[:300]}",
                      flush=True)
    print(f"{ran} cases of the GPU test file run through the REAL host code on the emulated runtime: {ran - failed} pass, {failed} fail")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
