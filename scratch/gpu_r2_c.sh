#!/usr/bin/env bash
# Round 2 call C (N = 1): new potrf driver (lookahead depth L, diagonal-first chain): parity, then the chain-bound proxy
# (small n on one GPU: the chain, not the trailing update, bounds the run -- the regime of 8 GPUs at n = 65536).
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2c_timeline.txt; }
for L in 1 2 3; do
  SB200_LOOKAHEAD=$L SB200_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "potrf or posv or mixed or stream or probe" > $OUT/r2c_pytest_L$L.log 2>&1; tail -3 $OUT/r2c_pytest_L$L.log; stamp pytest_L$L
done
for n in 8192 16384 32768; do for L in 1 2 3; do
  SB200_LOOKAHEAD=$L SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py potrf $n 512 2>> $OUT/r2c_perf_potrf.err | grep routine | sed "s/^/L=$L /" | cut -c1-200 | tee -a $OUT/r2c_perf_potrf.log
done; done; stamp perf
grep sb200_phases $OUT/r2c_perf_potrf.err | cut -c1-400
timeout 400 python bench.py --size 32768 --no-also --steps 5 --warmup 3 > $OUT/r2c_bench_potrf_n32768.json 2> $OUT/r2c_bench_potrf_n32768.err; cut -c1-700 $OUT/r2c_bench_potrf_n32768.json; tail -3 $OUT/r2c_bench_potrf_n32768.err; stamp bench
for sw in SB200_DIAG_RSQRT SB200_DIAG_WARP; do
  env $sw=1 timeout 300 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_kernels.py tests/test_reference_tester_gpu.py -m gpu -q --timeout 120 -n 4 -k "potrf or posv or trsm" > $OUT/r2c_switch_$sw.log 2>&1; tail -3 $OUT/r2c_switch_$sw.log; stamp $sw
done
