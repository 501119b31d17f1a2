#!/usr/bin/env bash
# Round 2 call L (N = 1): persistent multi-block LU base kernel (v5): parity of every variant, then timing per NBLK
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2l_timeline.txt; }
timeout 900 python -m pytest tests/test_zzz_gpu_round2_candidates.py -m gpu -q -x --timeout 300 -n 4 -k "base_kernel_variants" > $OUT/r2l_pytest.log 2>&1; tail -4 $OUT/r2l_pytest.log; stamp pytest_variants
for nblk in 1 2 4 8; do
  SB200_PANEL_NBLK=$nblk SB200_PHASES=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py getrf 32768 512 2>> $OUT/r2l_perf_getrf.err | grep routine | sed "s/^/NBLK=$nblk /" | cut -c1-200 | tee -a $OUT/r2l_perf_getrf.log
done; stamp perf_getrf
grep sb200_phases $OUT/r2l_perf_getrf.err | cut -c1-400
for nblk in 1 4 8; do
  SB200_PANEL_NBLK=$nblk timeout 300 python scratch/bench_contended.py 512 2> $OUT/r2l_contended_$nblk.err | grep "LU panel" | sed "s/^/NBLK=$nblk /"; grep sb200_phases $OUT/r2l_contended_$nblk.err | sed -n '1p;8p' | cut -c1-300
done; stamp contended
