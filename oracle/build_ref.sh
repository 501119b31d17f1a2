#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (SLATE + blaspp + lapackpp + matgen +
# testsweeper + tester) from the sources where they lie under /root/reference
# into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).  The
# reference's own build system is not used: g++ on the source files directly,
# a one-rank mpi.h stub (oracle/mpi_stub), and the OpenBLAS bundled with scipy
# as host BLAS/LAPACK.  CPU only (Target::HostTask); the device entry points
# resolve to the reference's own src/omptarget stubs.
#
# Products:
#   oracle/_ref/libslate_ref.so   reference library (slate+blaspp+lapackpp+matgen)
#   oracle/_ref/tester            reference tester (CPU baseline, residual checks)
#   oracle/_ref/ref_dump          oracle/ref_dump.cc linked to the library: dumps
#                                 Philox inputs and HostTask outputs for parity
# No reference source is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${SB200_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"
PY="${PYTHON:-python}"

if [ ! -d "$R/src/internal" ]; then
    echo "build_ref: $R not present; keeping whatever is prebuilt in $OUT" >&2
    exit 0
fi
OB="$($PY - <<'PYEOF'
import glob, os, scipy
d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
print(sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))[0])
PYEOF
)"
echo "build_ref: host BLAS = $OB"
mkdir -p "$OBJ"/{blaspp,lapackpp,slate,ts,test}

CXX="g++ -std=c++17 -fopenmp -fPIC -w"
I_BL="-I$HERE/cfg -I$R/blaspp/include -I$R/lapackpp/include"
I_SL="-I$HERE/mpi_stub $I_BL -I$R/include -I$R/src -I$R/matgen"

# compile_set <objdir> <flags...> -- <sources...>: parallel, incremental
compile_set() {
    local dir="$1"; shift
    local flags=()
    while [ "$1" != "--" ]; do flags+=("$1"); shift; done
    shift
    printf '%s\n' "$@" | xargs -P "$JOBS" -I{} bash -c '
        src="$1"; dir="$2"; shift 2
        o="$dir/$(echo "$src" | sed "s#/#_#g").o"
        if [ ! -f "$o" ] || [ "$src" -nt "$o" ]; then "$@" -c "$src" -o "$o" || exit 255; fi
    ' _ {} "$dir" $CXX "${flags[@]}"
}

shopt -s nullglob
BLASPP=( $(ls "$R"/blaspp/src/*.cc | grep -v -E '(cublas|rocblas|onemkl)_wrappers\.cc') )
LAPACKPP=( "$R"/lapackpp/src/*.cc "$R"/lapackpp/src/stub/*.cc )
SLATE=( "$R"/src/*.cc "$R"/src/internal/*.cc "$R"/src/work/*.cc "$R"/src/core/*.cc
        "$R"/src/auxiliary/*.cc "$R"/src/omptarget/*.cc "$R"/matgen/*.cc )
TS=( "$R"/testsweeper/testsweeper.cc "$R"/testsweeper/version.cc )
TEST=( "$R"/test/*.cc )

t0=$SECONDS
compile_set "$OBJ/blaspp"   -O2 $I_BL -- "${BLASPP[@]}";   echo "build_ref: blaspp   done ($((SECONDS-t0)) s)"
compile_set "$OBJ/lapackpp" -O1 $I_BL -- "${LAPACKPP[@]}"; echo "build_ref: lapackpp done ($((SECONDS-t0)) s)"
compile_set "$OBJ/slate"    -O2 $I_SL -- "${SLATE[@]}";    echo "build_ref: slate    done ($((SECONDS-t0)) s)"

g++ -shared -fopenmp -o "$OUT/libslate_ref.so" "$OBJ"/slate/*.o "$OBJ"/lapackpp/*.o "$OBJ"/blaspp/*.o \
    "$OB" -Wl,-rpath,"$(dirname "$OB")"
echo "build_ref: libslate_ref.so linked ($((SECONDS-t0)) s)"

$CXX -O2 $I_SL "$HERE/ref_dump.cc" -o "$OUT/ref_dump" \
    -L"$OUT" -lslate_ref "$OB" -Wl,-rpath,"$OUT" -Wl,-rpath,"$(dirname "$OB")" -Wl,-rpath,'$ORIGIN'
echo "build_ref: ref_dump linked ($((SECONDS-t0)) s)"

if [ "${SB200_SKIP_TESTER:-0}" != "1" ]; then
    compile_set "$OBJ/ts"   -O2 $I_SL -I"$R/testsweeper" -- "${TS[@]}"
    compile_set "$OBJ/test" -O1 $I_SL -I"$R/testsweeper" -I"$R/test" -- "${TEST[@]}"
    g++ -fopenmp -o "$OUT/tester" "$OBJ"/test/*.o "$OBJ"/ts/*.o \
        -L"$OUT" -lslate_ref "$OB" -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$(dirname "$OB")"
    echo "build_ref: tester linked ($((SECONDS-t0)) s)"
fi
echo "build_ref: OK"
