#!/usr/bin/env bash
# Round 2 call A (N = 1): guarded tests, kernel-level chain timings, a trimmed variant sweep for potrf / getrf.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
bash scratch/gpu_r2_call1.sh tests
timeout 300 python scratch/bench_tile.py 512 32 > $OUT/r2a_bench_tile.log 2> $OUT/r2a_bench_tile.err; cut -c1-200 $OUT/r2a_bench_tile.log
SB200_VARIANTS="default,diag_mw,diag_mw+trsm_fused,tile+trsm_fused" timeout 400 python scratch/perf_variants.py potrf 32768 512 > $OUT/r2a_perf_potrf.log 2> $OUT/r2a_perf_potrf.err
cut -c1-400 $OUT/r2a_perf_potrf.log
SB200_VARIANTS="default,diag_mw,skinny_panel_update,transposed_U_row,panel_ll+all_row_solves+diag_mw,everything" timeout 500 python scratch/perf_variants.py getrf 32768 512 > $OUT/r2a_perf_getrf.log 2> $OUT/r2a_perf_getrf.err
cut -c1-400 $OUT/r2a_perf_getrf.log
timeout 200 python scratch/perf_variants.py gemm 16384 512 > $OUT/r2a_perf_gemm.log 2> $OUT/r2a_perf_gemm.err; cut -c1-300 $OUT/r2a_perf_gemm.log
