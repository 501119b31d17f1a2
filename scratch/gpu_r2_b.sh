#!/usr/bin/env bash
# Round 2 call B (N = 1): new defaults through the whole GPU suite, the new bench line, contended chain kernels,
# e2e modes, the drop-in testers (cuBLAS incumbent vs our kernels), diag switches.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2b_timeline.txt; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 > $OUT/r2b_pytest.log 2>&1; tail -5 $OUT/r2b_pytest.log; stamp pytest
SB200_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_zzz_gpu_round2_candidates.py tests/test_zzz_gpu_dist_solve.py -m gpu -q --timeout 120 -n 4 > $OUT/r2b_pytest_guarded.log 2>&1; tail -5 $OUT/r2b_pytest_guarded.log; stamp guarded
timeout 300 python scratch/bench_contended.py 512 > $OUT/r2b_contended.log 2> $OUT/r2b_contended.err; cat $OUT/r2b_contended.log; grep sb200_phases $OUT/r2b_contended.err | cut -c1-400; stamp contended
for sw in SB200_DIAG_RSQRT SB200_DIAG_WARP; do env $sw=1 timeout 200 python scratch/bench_tile.py 512 32 2>&1 | grep -E "potrf_tile_d.*default" | sed "s/^/$sw /"; done; stamp diag
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r2b_bench_default.json 2> $OUT/r2b_bench_default.err; cut -c1-2500 $OUT/r2b_bench_default.json; tail -3 $OUT/r2b_bench_default.err; stamp bench_default
for m in 0 1 2; do
  SB200_E2E_OVERLAP=$m timeout 300 python bench.py --size 32768 --no-also --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2b_bench_e2e$m.json 2> $OUT/r2b_bench_e2e$m.err
  grep -o '"e2e": {[^}]*}' $OUT/r2b_bench_e2e$m.json | cut -c1-300; grep -o '"value": [0-9.]*' $OUT/r2b_bench_e2e$m.json | head -1
done; stamp e2e
export OMP_NUM_THREADS=16 OPENBLAS_NUM_THREADS=1
for t in tester_cublas tester_sb200; do
  for r in potrf getrf gemm; do
    timeout 400 oracle/_ref/$t --target d --origin d --type d --dim 16384 --nb 512 --check n --ref n --repeat 3 $r > $OUT/r2b_${t}_$r.log 2>&1
    grep -E "^ +d|pass|FAIL|error" $OUT/r2b_${t}_$r.log | tail -4 | cut -c1-260
  done
done; stamp testers
