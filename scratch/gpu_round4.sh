#!/usr/bin/env bash
# Third N = 1 call: re-validate the three re-tuned kernels (trsm_small, skinny GEMM, norms) and re-measure.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/timeline4.txt; }
stamp start
timeout 300 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new4.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_new4.log
tail -12 $OUT/pytest_new4.log | cut -c1-300
stamp new_tests
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_reference_tester_gpu.py tests/test_gpu_drivers.py tests/test_gpu_mixed.py -m gpu -q --timeout 300 -n 6 > $OUT/pytest_gpu4.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu4.log
tail -8 $OUT/pytest_gpu4.log | cut -c1-300
stamp old_tests
B="timeout 300 python bench.py"
$B --routine posv_mixed --steps 3 > $OUT/bench4_posv_mixed.json 2> $OUT/bench4_posv_mixed.err; tail -1 $OUT/bench4_posv_mixed.json | cut -c1-250; grep -o '"step_ms.*' $OUT/bench4_posv_mixed.json; tail -3 $OUT/bench4_posv_mixed.err
$B --routine gesv_mixed --steps 3 > $OUT/bench4_gesv_mixed.json 2> $OUT/bench4_gesv_mixed.err; tail -1 $OUT/bench4_gesv_mixed.json | cut -c1-250; grep -o '"step_ms.*' $OUT/bench4_gesv_mixed.json; tail -3 $OUT/bench4_gesv_mixed.err
stamp bench_mixed
$B --routine tileops --steps 5 > $OUT/bench4_tileops.json 2> $OUT/bench4_tileops.err; grep -o '"genorm_max.*' $OUT/bench4_tileops.json | cut -c1-900; tail -3 $OUT/bench4_tileops.err
stamp bench_tileops
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches4_posv_mixed.csv \
    python scratch/prof_mixed.py 8192 > $OUT/ncu4_launches.log 2>&1
stamp ncu_launches
timeout 200 ncu --set full --clock-control none --import-source on -k regex:norm_kernel -s 2 -c 2 -f -o $OUT/prof4_norm \
    python scratch/prof_norm.py > $OUT/ncu4_norm.log 2>&1
stamp ncu_norm
