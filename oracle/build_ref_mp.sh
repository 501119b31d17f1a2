#!/usr/bin/env bash
# oracle/build_ref_mp.sh -- TEST INFRASTRUCTURE ONLY.
# The unmodified reference once more, against the MULTI-PROCESS MPI replacement (oracle/mpi_mp) instead of the one-rank stub,
# so that its Target::HostTask drivers run on a p x q process grid here (no MPI in the image):
#   oracle/_ref/libslate_ref_mp.so   SLATE + matgen compiled with oracle/mpi_mp/mpi.h (blaspp / lapackpp objects: those of
#                                    oracle/build_ref.sh, they do not see MPI)
#   oracle/_ref/ref_dump_mp          oracle/ref_dump.cc linked to it; run it through  python oracle/mprun.py -n N ... p=P q=Q
#   oracle/_ref/tester_mp            the reference's own tester:  python oracle/mprun.py -n 4 oracle/_ref/tester_mp --grid 2x2 ... getrf
# Sources are compiled where they lie under /root/reference; nothing is copied.  Needs oracle/build_ref.sh to have run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${SB200_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
OBJMP="$OUT/obj_mp"
JOBS="${JOBS:-$(nproc)}"
PY="${PYTHON:-python}"
if [ ! -d "$R/src/internal" ]; then
    echo "build_ref_mp: $R not present; keeping whatever is prebuilt in $OUT" >&2
    exit 0
fi
if [ ! -d "$OBJ/blaspp" ] || [ ! -d "$OBJ/lapackpp" ]; then
    echo "build_ref_mp: run oracle/build_ref.sh first (blaspp / lapackpp objects)" >&2
    exit 1
fi
OB="$($PY - <<'PYEOF'
import glob, os, scipy
d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
print(sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))[0])
PYEOF
)"
mkdir -p "$OBJMP/slate"
CXX="g++ -std=c++17 -fopenmp -fPIC -w"
I_BL="-I$HERE/cfg -I$R/blaspp/include -I$R/lapackpp/include"
I_SL="-I$HERE/mpi_mp $I_BL -I$R/include -I$R/src -I$R/matgen"
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
SLATE=( "$R"/src/*.cc "$R"/src/internal/*.cc "$R"/src/work/*.cc "$R"/src/core/*.cc
        "$R"/src/auxiliary/*.cc "$R"/src/omptarget/*.cc "$R"/matgen/*.cc )
t0=$SECONDS
compile_set "$OBJMP/slate" -O2 $I_SL -- "${SLATE[@]}"; echo "build_ref_mp: slate done ($((SECONDS-t0)) s)"
$CXX -O2 -I"$HERE/mpi_mp" -c "$HERE/mpi_mp/mpi_mp.cc" -o "$OBJMP/mpi_mp.o"
g++ -shared -fopenmp -o "$OUT/libslate_ref_mp.so" "$OBJMP"/slate/*.o "$OBJMP/mpi_mp.o" "$OBJ"/lapackpp/*.o "$OBJ"/blaspp/*.o \
    "$OB" -Wl,-rpath,"$(dirname "$OB")" -lpthread
$CXX -O2 $I_SL "$HERE/ref_dump.cc" -o "$OUT/ref_dump_mp" \
    -L"$OUT" -lslate_ref_mp "$OB" -Wl,-rpath,"$OUT" -Wl,-rpath,"$(dirname "$OB")" -Wl,-rpath,'$ORIGIN' -lpthread
echo "build_ref_mp: ref_dump_mp linked ($((SECONDS-t0)) s)"

# the reference's own tester on process grids: python oracle/mprun.py -n 4 oracle/_ref/tester_mp --grid 2x2 ... getrf
if [ "${SB200_SKIP_TESTER:-0}" != "1" ]; then
    mkdir -p "$OBJMP/ts" "$OBJMP/test"
    TS=( "$R"/testsweeper/testsweeper.cc "$R"/testsweeper/version.cc )
    TEST=( "$R"/test/*.cc )
    compile_set "$OBJMP/ts"   -O2 $I_SL -I"$R/testsweeper" -- "${TS[@]}"
    compile_set "$OBJMP/test" -O1 $I_SL -I"$R/testsweeper" -I"$R/test" -- "${TEST[@]}"
    g++ -fopenmp -o "$OUT/tester_mp" "$OBJMP"/test/*.o "$OBJMP"/ts/*.o \
        -L"$OUT" -lslate_ref_mp "$OB" -Wl,-rpath,"$OUT" -Wl,-rpath,"$(dirname "$OB")" -Wl,-rpath,'$ORIGIN' -lpthread
    echo "build_ref_mp: tester_mp linked ($((SECONDS-t0)) s)"
fi
