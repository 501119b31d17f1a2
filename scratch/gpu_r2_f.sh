#!/usr/bin/env bash
# v3 LU base kernel: per-phase cycle trace (debug build switch SB200_V3_TRACE=1)
OUT=gpurun_out; mkdir -p $OUT
SB200_V3_TRACE=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py getrf 16384 512 2> $OUT/r2f_trace.err | grep routine | cut -c1-200
grep v3_trace $OUT/r2f_trace.err | head -60
