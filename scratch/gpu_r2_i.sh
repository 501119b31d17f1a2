#!/usr/bin/env bash
# Round 2 call I (N = 1): register-resident LU base kernel (v4): parity, phases, panel timing
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2i_timeline.txt; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "getrf or gesv or lu or panel or permute" > $OUT/r2i_pytest_lu.log 2>&1; tail -3 $OUT/r2i_pytest_lu.log; stamp pytest_lu
for n in 16384 32768; do
  SB200_PHASES=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py getrf $n 512 2>> $OUT/r2i_perf_getrf.err | grep routine | cut -c1-200 | tee -a $OUT/r2i_perf_getrf.log
done; stamp perf_getrf
grep sb200_phases $OUT/r2i_perf_getrf.err | cut -c1-400
timeout 300 python scratch/bench_contended.py 512 > $OUT/r2i_contended.log 2> $OUT/r2i_contended.err; grep "LU panel" $OUT/r2i_contended.log; grep sb200_phases $OUT/r2i_contended.err | head -4 | cut -c1-300; stamp contended
