#!/usr/bin/env bash
# Final N = 1 call of the round: the rewritten small-nrhs solve kernel (tests), mixed bench lines, launch list, big tcgen05 capture.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/timeline5.txt; }
stamp start
timeout 300 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new5.log 2>&1; rc=$?; echo "pytest exit $rc" >> $OUT/pytest_new5.log
tail -12 $OUT/pytest_new5.log | cut -c1-300
stamp new_tests
B="timeout 300 python bench.py"
$B --routine posv_mixed --steps 3 > $OUT/bench5_posv_mixed.json 2> $OUT/bench5_posv_mixed.err; tail -1 $OUT/bench5_posv_mixed.json | cut -c1-250; grep -o '"step_ms.*' $OUT/bench5_posv_mixed.json; tail -3 $OUT/bench5_posv_mixed.err
$B --routine gesv_mixed --steps 3 > $OUT/bench5_gesv_mixed.json 2> $OUT/bench5_gesv_mixed.err; tail -1 $OUT/bench5_gesv_mixed.json | cut -c1-250; grep -o '"step_ms.*' $OUT/bench5_gesv_mixed.json; tail -3 $OUT/bench5_gesv_mixed.err
stamp bench_mixed
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches5_posv_mixed.csv \
    python scratch/prof_mixed.py 8192 > $OUT/ncu5_launches.log 2>&1
stamp ncu_launches
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f -o $OUT/prof5_tf32x3_big \
    python scratch/prof_mixed.py 16384 > $OUT/ncu5_tf32x3.log 2>&1
stamp ncu_tf32x3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke5.log 2>&1; tail -2 $OUT/smoke5.log
stamp smoke
