#!/usr/bin/env bash
# Round 2, FIRST GPU call(s) (N = 1): validate everything written after round 1's GPU budget was spent, then measure
# every candidate against the default path.  Nothing here changes defaults.  Sections (box time, roughly):
#   tests 6 min | switches 4 min | perf 10 min | bench 4 min | ncu 3 min        all = ~27 min
#   gpurun --timeout 1800 -- 'bash scratch/gpu_r2_call1.sh all'      or one section per call, e.g.  ... gpu_r2_call1.sh tests
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2c1_timeline.txt; }
stamp start
sec_tests() {
  # 1. smoke + the guarded candidate tests (each file separately so that one failure does not hide the others)
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2c1_smoke.log 2>&1; tail -2 $OUT/r2c1_smoke.log
  stamp smoke
  # groups in order of increasing risk of a hang (spin-wait protocols last); --timeout-method=thread makes pytest-timeout
  # os._exit() a worker that is stuck inside a CUDA call, and the outer timeout -k kills whatever is left
  PT="python -m pytest -m gpu -q --timeout 90 --timeout-method=thread -n 4"
  C=tests/test_zzz_gpu_round2_candidates.py
  run_group() {   # name, pytest args...
    local name=$1; shift
    SB200_RUN_UNVALIDATED=1 timeout -k 10 400 $PT "$@" > $OUT/r2c1_$name.log 2>&1
    echo "pytest exit $?" >> $OUT/r2c1_$name.log; tail -12 $OUT/r2c1_$name.log | cut -c1-240
    stamp $name
  }
  run_group dist_solve tests/test_zzz_gpu_dist_solve.py
  run_group diag_mw    $C -k "diag_mw"
  run_group streaming  $C -k "streaming or permute_rows or transposed or skinny_panel or her2k or scalapack or nopiv or symmetric_rank or symm_matches or trmm"
  run_group trsm_fused $C -k "fused_panel_trsm or fused_row_trsm or fused_row_solve or row_solve_candidates"
  run_group tile_fused $C -k "fused_tile"
  run_group panel_ll   $C -k "ll_panel"
  run_group all_fused  $C -k "all_fused or tile_and_panel_solve"
  nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader | tee -a $OUT/r2c1_timeline.txt   # nothing may be left running
}
sec_switches() {
  # 2. latched switches: the existing parity tests under each of them (fresh process per switch)
  for sw in SB200_DIAG_RSQRT SB200_DIAG_WARP SB200_PANEL_BARRIER SB200_PANEL_LL; do
    env $sw=1 timeout -k 10 300 python -m pytest tests/test_zz_gpu_panel_variants.py tests/test_gpu_drivers.py -m gpu -q --timeout 90 --timeout-method=thread -n 4 \
        > $OUT/r2c1_switch_$sw.log 2>&1
    echo "pytest exit $?" >> $OUT/r2c1_switch_$sw.log; tail -6 $OUT/r2c1_switch_$sw.log | cut -c1-240
    stamp $sw
  done
}
sec_perf() {
  # 3a. kernel-level timings of the chain pieces (potrf tile, panel / row / small solves), every per-call switch
  timeout 300 python scratch/bench_tile.py 512 32 > $OUT/r2c1_bench_tile.log 2> $OUT/r2c1_bench_tile.err; cat $OUT/r2c1_bench_tile.log | cut -c1-200
  stamp bench_tile
  # 3. timings of every variant (one fresh process each), phases on stderr
  timeout 300 python scratch/perf_variants.py gemm 16384 512 > $OUT/r2c1_perf_gemm.log 2> $OUT/r2c1_perf_gemm.err; cat $OUT/r2c1_perf_gemm.log | cut -c1-300
  stamp perf_gemm
  for r in potrf getrf posv_mixed gesv_mixed; do
    timeout 900 python scratch/perf_variants.py $r 32768 512 > $OUT/r2c1_perf_$r.log 2> $OUT/r2c1_perf_$r.err
    cat $OUT/r2c1_perf_$r.log | cut -c1-300
    stamp perf_$r
  done
}
sec_bench() {
  # 4. e2e of the default bench with the overlap modes (0 = none, 1 = D2H streamed, 2 = H2D + D2H streamed)
  for m in 0 1 2; do
    SB200_E2E_OVERLAP=$m timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/r2c1_bench_e2e$m.json 2> $OUT/r2c1_bench_e2e$m.err
    grep -o '"e2e": {[^}]*}' $OUT/r2c1_bench_e2e$m.json | cut -c1-200
    stamp bench_e2e$m
  done
  # 5. N = 1 at the metric's size (n = 65536 fits one B200: 16 GiB of lower tiles)
  timeout 400 python bench.py --routine potrf --size 65536 --steps 2 --warmup 3 --no-e2e > $OUT/r2c1_bench_potrf_n65536.json 2> $OUT/r2c1_bench_potrf_n65536.err
  tail -1 $OUT/r2c1_bench_potrf_n65536.json | cut -c1-300; tail -3 $OUT/r2c1_bench_potrf_n65536.err
  stamp bench_n65536
}
sec_ncu() {
  # 6. refreshed launch list of the default bench command (profiles/r01_launches_potrf_summary.txt predates the fast diagonal kernels)
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2c1_launches_potrf.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --size 8192 > $OUT/r2c1_ncu_launches.log 2>&1
  stamp ncu_launches
  # 7. per-launch durations of the candidates themselves (one pass, no replay): potrf n=2048 (4 diagonal tiles) with the fused
  #    tile + panel solve, getrf n=4096 with the LL panel + fused row solves
  SB200_DIAG_MW=1 SB200_TILE_FUSED=1 SB200_TRSM_FUSED=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/r2c1_launches_potrf2048_fused.csv python scratch/prof_potrf_small.py > $OUT/r2c1_ncu_fused.log 2>&1
  SB200_DIAG_MW=1 SB200_PANEL_LL=1 SB200_TRSM_FUSED=6 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/r2c1_launches_getrf4096_ll.csv python scratch/prof_getrf_small.py > $OUT/r2c1_ncu_ll.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/r2c1_launches_getrf4096_default.csv python scratch/prof_getrf_small.py > $OUT/r2c1_ncu_getrf_default.log 2>&1
  python scratch/launch_summary.py $OUT/r2c1_launches_potrf.csv $OUT/r2c1_launches_potrf2048_fused.csv $OUT/r2c1_launches_getrf4096_ll.csv \
      $OUT/r2c1_launches_getrf4096_default.csv > $OUT/r2c1_launch_summaries.txt 2>&1; head -60 $OUT/r2c1_launch_summaries.txt
  stamp ncu_candidates
}
WHAT=${1:-all}
case $WHAT in
  all) sec_tests; sec_switches; sec_perf; sec_bench; sec_ncu ;;
  tests|switches|perf|bench|ncu) sec_$WHAT ;;
  *) echo "usage: $0 [all|tests|switches|perf|bench|ncu]"; exit 2 ;;
esac
stamp done
