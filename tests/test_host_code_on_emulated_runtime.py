"""The SHIPPED host code of the matrix-level BLAS-3 / solve widening (slate_b200/csrc/solve.cu: trmm / hemm / symm variants,
gemm and the rank-k updates with transposed views, the right-side triangular sweep, getrs with an op, the norms' host
combination) executed on the CPU, where there is no GPU: scratch/cpu_standin/fake_cudart.cc is LD_PRELOADed in place of
libcudart -- device memory is host memory, every kernel those drivers launch is replaced by a plain loop that computes what
the kernel is specified to compute, a kernel without an emulation fails the launch loudly -- and the cases of
tests/test_zzzzz_gpu_blas3_variants.py (here: every shape up to 300) run through slate_b200.host -> C ABI -> solve.cu.

This is a test of HOST LOGIC (plan building, tile indices, operand roles, launch order, in-place updates); it says nothing
about the kernels, which only a B200 can run, and it is not a CPU path of the product: the library is unchanged and cannot
find this runtime by itself."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


def test_new_drivers_host_code_passes_their_gpu_tests_on_the_emulated_runtime(tmp_path):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime_api.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    lib = str(tmp_path / "libfakecudart.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + CUDA_INC,
                    os.path.join(ROOT, "scratch", "cpu_standin", "fake_cudart.cc"), "-o", lib], check=True, timeout=300)
    env = dict(os.environ, LD_PRELOAD=lib, STANDIN_MAX_DIM="300")
    env.pop("STANDIN_ALL", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scratch", "cpu_standin", "run_real_host_code_on_cpu.py")],
                       capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    text = r.stdout + r.stderr
    m = re.search(r"(\d+) cases .* emulated runtime: (\d+) pass, (\d+) fail", text)
    assert m, text[-3000:]
    assert "NO EMULATION" not in text, text[-3000:]
    assert int(m.group(3)) == 0 and r.returncode == 0, text[-3000:]
    assert int(m.group(1)) >= 300                        # trmm / hemm / symm / trsm / gemm / rank-k / getrs / norm cases
