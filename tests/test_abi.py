"""CPU tests of the drop-in boundary: libslate_b200.so loads without a GPU and exports every
entry point include/slate_b200.h declares; argument errors come back as codes, not crashes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slate_b200.h")


def declared_symbols():
    """Every function the header declares, after the C preprocessor has expanded the per-type
    X-macros (SB200_FOR_TYPES)."""
    import subprocess
    src = subprocess.run(["gcc", "-E", "-P", "-x", "c", HEADER], capture_output=True, text=True, check=True).stdout
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_surface():
    syms = declared_symbols()
    assert len(syms) > 150          # 4 scalar types x every seam-1 / seam-2 entry point
    for must in ["sb200_gemm_batched_z", "sb200_trsm_batched_c", "sb200_potrf_tile_z", "sb200_gecopy_batched_dz",
                 "sb200_henorm_batched_c", "sb200_transpose_inplace_s", "sb200_gemm_strided_d",
                 "sb200_gemm_batched_d", "sb200_herk_batched_d", "sb200_trsm_batched_d", "sb200_potrf_tile_d",
                 "sb200_permute_rows_d", "sb200_geadd_batched_d", "sb200_genorm_batched_d",
                 "sb200_transpose_batched_d", "sb200_potrf_d", "sb200_getrf_d", "sb200_gemm_d"]:
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol():
    from slate_b200._lib import lib, LIB_PATH
    assert os.path.exists(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/slate_b200.h but not exported: {missing}"
    assert lib.sb200_version() >= 100


def test_error_codes_without_compute():
    from slate_b200._lib import lib
    lib.sb200_strerror.restype = ctypes.c_char_p
    assert b"invalid" in lib.sb200_strerror(-1)
    assert b"no CUDA device" in lib.sb200_strerror(-4)
    i64, dbl, ptr = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    f = lib.sb200_gemm_batched_d
    f.argtypes = [ctypes.c_int] * 3 + [i64] * 3 + [dbl, ptr, i64, ptr, i64, dbl, ptr, i64, i64, ptr]
    # bad op code -> EINVAL; zero batch -> quick return (reference: device_geadd.cu:127-129)
    assert f(ord("C"), ord("X"), ord("N"), 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, 1, None) == -1
    assert f(ord("C"), ord("N"), ord("N"), 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, 0, None) == 0
    assert f(ord("C"), ord("N"), ord("N"), 4, 4, 4, 1.0, None, 2, None, 4, 0.0, None, 4, 1, None) == -1   # lda < m
    g = lib.sb200_trsm_batched_d
    g.argtypes = [ctypes.c_int] * 5 + [i64, i64, dbl, ptr, i64, ptr, i64, i64, ptr, ptr]
    assert g(ord("C"), ord("Q"), ord("L"), ord("N"), ord("N"), 4, 4, 1.0, None, 4, None, 4, 1, None, None) == -1


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle (tests/bench/smoke only)."""
    pkg = os.path.join(ROOT, "slate_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "slate_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn


def test_grid_shape_choice_matches_reference_rule():
    # as square as possible, p <= q (reference tester: test/test.cc:738-747)
    from slate_b200.host import Grid
    assert Grid.choose(1) == (1, 1)
    assert Grid.choose(2) == (1, 2)
    assert Grid.choose(4) == (2, 2)
    assert Grid.choose(8) == (2, 4)
    assert Grid.choose(6) == (2, 3)
