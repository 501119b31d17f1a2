#!/usr/bin/env bash
# oracle/build_ref_gpu.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference with its CUDA device support (-DBLAS_HAVE_CUBLAS) from the
# sources under /root/reference, twice linked:
#   oracle/_ref/tester_cublas  stock reference: blaspp -> cuBLAS, lapackpp -> cuSOLVER, src/cuda/*.cu
#                              (the GPU incumbent)
#   oracle/_ref/tester_sb200   the SAME objects, except that the files listed in DROPPED below are
#                              left out and shim/*.cc + libslate_b200.so take their place.  This is
#                              the drop-in proof: the reference's own tester (Target::Devices) and its
#                              residual checks run on our kernels without touching a reference source.
# The reference's own build system is not used; no reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
R="${SB200_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj/gpu"
JOBS="${JOBS:-$(nproc)}"
PY="${PYTHON:-python}"
CUDA="${CUDA_HOME:-/usr/local/cuda}"

if [ ! -d "$R/src/internal" ]; then
    echo "build_ref_gpu: $R not present; keeping whatever is prebuilt in $OUT" >&2
    exit 0
fi
OB="$($PY - <<'PYEOF'
import glob, os, scipy
d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
print(sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))[0])
PYEOF
)"
mkdir -p "$OBJ"/{blaspp,lapackpp,slate,cuda,ts,test,shim}

DEFS="-DBLAS_HAVE_CUBLAS -DLAPACK_HAVE_CUBLAS"
CXX="g++ -std=c++17 -fopenmp -fPIC -w $DEFS"
I_BL="-I$HERE/cfg -I$R/blaspp/include -I$R/lapackpp/include -I$CUDA/include"
I_SL="-I$HERE/mpi_stub $I_BL -I$R/include -I$R/src -I$R/matgen"

compile_set() {     # <objdir> <compiler words...> -- <sources...>: parallel, incremental
    local dir="$1"; shift
    local cc=()
    while [ "$1" != "--" ]; do cc+=("$1"); shift; done
    shift
    printf '%s\n' "$@" | xargs -P "$JOBS" -I{} bash -c '
        src="$1"; dir="$2"; shift 2
        o="$dir/$(echo "$src" | sed "s#/#_#g").o"
        if [ ! -f "$o" ] || [ "$src" -nt "$o" ]; then "$@" -c "$src" -o "$o" || exit 255; fi
    ' _ {} "$dir" "${cc[@]}"
}

shopt -s nullglob
BLASPP=( $(ls "$R"/blaspp/src/*.cc | grep -v -E '(rocblas|onemkl)_wrappers\.cc') )
LAPACKPP=( "$R"/lapackpp/src/*.cc "$R"/lapackpp/src/cuda/*.cc )
SLATE=( "$R"/src/*.cc "$R"/src/internal/*.cc "$R"/src/work/*.cc "$R"/src/core/*.cc
        "$R"/src/auxiliary/*.cc "$R"/matgen/*.cc )
SLATE_CU=( "$R"/src/cuda/*.cu )
BLASPP_CU=( "$R"/blaspp/src/cuda/*.cu )
TS=( "$R"/testsweeper/testsweeper.cc "$R"/testsweeper/version.cc )
TEST=( "$R"/test/*.cc )
SHIM=( "$ROOT"/shim/blaspp_shim.cc "$ROOT"/shim/lapackpp_shim.cc "$ROOT"/shim/slate_device_shim.cc )

t0=$SECONDS
compile_set "$OBJ/blaspp"   $CXX -O2 $I_BL -- "${BLASPP[@]}";   echo "build_ref_gpu: blaspp   done ($((SECONDS-t0)) s)"
compile_set "$OBJ/lapackpp" $CXX -O1 $I_BL -- "${LAPACKPP[@]}"; echo "build_ref_gpu: lapackpp done ($((SECONDS-t0)) s)"
compile_set "$OBJ/slate"    $CXX -O2 $I_SL -- "${SLATE[@]}";    echo "build_ref_gpu: slate    done ($((SECONDS-t0)) s)"
compile_set "$OBJ/cuda" nvcc -std=c++17 -O2 -w -gencode arch=compute_100a,code=sm_100a $DEFS \
    -Xcompiler -fPIC,-fopenmp $I_SL -- "${SLATE_CU[@]}";        echo "build_ref_gpu: src/cuda done ($((SECONDS-t0)) s)"
compile_set "$OBJ/blaspp" nvcc -std=c++17 -O2 -w -gencode arch=compute_100a,code=sm_100a $DEFS \
    -Xcompiler -fPIC,-fopenmp $I_BL -- "${BLASPP_CU[@]}"
compile_set "$OBJ/ts"   $CXX -O2 $I_SL -I"$R/testsweeper" -- "${TS[@]}"
compile_set "$OBJ/test" $CXX -O1 $I_SL -I"$R/testsweeper" -I"$R/test" -- "${TEST[@]}"
echo "build_ref_gpu: tester objects done ($((SECONDS-t0)) s)"
compile_set "$OBJ/shim" $CXX -O2 $I_SL -I"$ROOT/include" -I"$ROOT/shim" -- "${SHIM[@]}"

CUDALIBS="-L$CUDA/lib64 -lcublas -lcusolver -lcudart"
RPATH="-Wl,-rpath,\$ORIGIN -Wl,-rpath,$(dirname "$OB") -Wl,-rpath,$CUDA/lib64"

# (a) stock GPU reference
g++ -shared -fopenmp -o "$OUT/libslate_ref_cublas.so" "$OBJ"/slate/*.o "$OBJ"/cuda/*.o "$OBJ"/lapackpp/*.o "$OBJ"/blaspp/*.o \
    "$OB" $CUDALIBS $RPATH
g++ -fopenmp -o "$OUT/tester_cublas" "$OBJ"/test/*.o "$OBJ"/ts/*.o -L"$OUT" -lslate_ref_cublas "$OB" $CUDALIBS $RPATH
echo "build_ref_gpu: tester_cublas linked ($((SECONDS-t0)) s)"

# (b) drop-in: leave out exactly the reference files our shim replaces
DROPPED='blaspp_src_device_batch_(gemm|herk|syrk|trsm)\.cc\.o|blaspp_src_device_(herk|syrk)\.cc\.o|lapackpp_src_cuda_cuda_potrf\.cc\.o'
KEEP=( $(ls "$OBJ"/slate/*.o "$OBJ"/lapackpp/*.o "$OBJ"/blaspp/*.o | grep -v -E "$DROPPED") )
mkdir -p "$OUT/lib"
cp -f "$ROOT/slate_b200/lib/libslate_b200.so" "$OUT/lib/" 2>/dev/null || true
g++ -shared -fopenmp -o "$OUT/libslate_ref_sb200.so" "${KEEP[@]}" "$OBJ"/shim/*.o \
    -L"$ROOT/slate_b200/lib" -lslate_b200 "$OB" $CUDALIBS $RPATH -Wl,-rpath,'$ORIGIN/../../slate_b200/lib'
g++ -fopenmp -o "$OUT/tester_sb200" "$OBJ"/test/*.o "$OBJ"/ts/*.o -L"$OUT" -lslate_ref_sb200 \
    -L"$ROOT/slate_b200/lib" -lslate_b200 "$OB" $CUDALIBS $RPATH -Wl,-rpath,'$ORIGIN/../../slate_b200/lib'
rm -rf "$OUT/lib"
echo "build_ref_gpu: tester_sb200 linked ($((SECONDS-t0)) s)"

# (c) the reference's own device unit tests (kernel vs host loops / lapack::lange / blas::gemm:
# unit_test/test_{geadd,gescale,geset,gecopy,norm,internal_blas}.cc), linked against the drop-in library
UNIT=( geadd gescale geset gecopy norm internal_blas )
mkdir -p "$OBJ/unit"
compile_set "$OBJ/unit" $CXX -O1 $I_SL -I"$R/testsweeper" -I"$R/unit_test" -- "$R/unit_test/unit_test.cc" \
    $(for u in "${UNIT[@]}"; do echo "$R/unit_test/test_$u.cc"; done)
for u in "${UNIT[@]}"; do
    g++ -fopenmp -o "$OUT/unit_${u}_sb200" "$OBJ"/unit/*unit_test_unit_test.cc.o "$OBJ"/unit/*unit_test_test_$u.cc.o \
        "$OBJ"/ts/*testsweeper.cc.o -L"$OUT" -lslate_ref_sb200 -L"$ROOT/slate_b200/lib" -lslate_b200 "$OB" $CUDALIBS $RPATH \
        -Wl,-rpath,'$ORIGIN/../../slate_b200/lib'
done
echo "build_ref_gpu: unit tests linked ($((SECONDS-t0)) s)"
echo "build_ref_gpu: OK"
