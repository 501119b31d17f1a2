#!/usr/bin/env bash
# Second N = 1 call of the session: validates the small-nrhs solve kernel, the skinny GEMM and the faster norm
# kernels (tests first, with the new kernels switched off one at a time if something fails), then re-measures.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/timeline3.txt; }
stamp start
timeout 420 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new3.log 2>&1; rc=$?; echo "pytest exit $rc" >> $OUT/pytest_new3.log
tail -30 $OUT/pytest_new3.log | cut -c1-300
stamp new_tests
if [ $rc -ne 0 ]; then
  SB200_SKINNY=0 timeout 300 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new3_noskinny.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_new3_noskinny.log
  tail -12 $OUT/pytest_new3_noskinny.log | cut -c1-300
  SB200_TRSM_SMALL=0 timeout 300 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new3_nosmall.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_new3_nosmall.log
  tail -12 $OUT/pytest_new3_nosmall.log | cut -c1-300
  stamp new_tests_isolation
fi
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -n 6 --deselect tests/test_zy_gpu_widening.py > $OUT/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu3.log
tail -15 $OUT/pytest_gpu3.log | cut -c1-300
stamp old_tests
B="timeout 300 python bench.py"
$B --routine posv_mixed --steps 2 > $OUT/bench3_posv_mixed.json 2> $OUT/bench3_posv_mixed.err; tail -1 $OUT/bench3_posv_mixed.json | cut -c1-300; grep -o '"phases_ms.*' $OUT/bench3_posv_mixed.json; tail -3 $OUT/bench3_posv_mixed.err
$B --routine gesv_mixed --steps 2 > $OUT/bench3_gesv_mixed.json 2> $OUT/bench3_gesv_mixed.err; tail -1 $OUT/bench3_gesv_mixed.json | cut -c1-300; grep -o '"phases_ms.*' $OUT/bench3_gesv_mixed.json; tail -3 $OUT/bench3_gesv_mixed.err
stamp bench_mixed
$B --routine tileops --steps 5 > $OUT/bench3_tileops.json 2> $OUT/bench3_tileops.err; grep -o '"kernels.*' $OUT/bench3_tileops.json | cut -c1-1800; tail -3 $OUT/bench3_tileops.err
stamp bench_tileops
$B > $OUT/bench3_potrf.json 2> $OUT/bench3_potrf.err; tail -1 $OUT/bench3_potrf.json | cut -c1-300
stamp bench_potrf
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches3_posv_mixed.csv \
    python scratch/prof_mixed.py 8192 > $OUT/ncu3_launches.log 2>&1
stamp ncu_launches
timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 0 -c 1 -f -o $OUT/prof3_tf32x3_big \
    python scratch/prof_mixed.py 16384 > $OUT/ncu3_tf32x3.log 2>&1
stamp ncu_tf32x3
timeout 240 ncu --set full --clock-control none --import-source on -k regex:norm_kernel -s 2 -c 2 -f -o $OUT/prof3_norm \
    python scratch/prof_norm.py > $OUT/ncu3_norm.log 2>&1
stamp ncu_norm
ls -la $OUT | tail -20
