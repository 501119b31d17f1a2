#!/usr/bin/env bash
# ONE gpurun call (N = 1), ordered by priority; every step has its own timeout and writes to gpurun_out/ as it goes.
# usage: gpurun --timeout 1300 -- 'bash scratch/gpu_round2.sh'
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/timeline.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
stamp start
# 1. the new tests first (small), then the whole existing GPU suite in parallel workers
timeout 420 python -m pytest tests/test_zy_gpu_widening.py -m gpu -q --timeout 180 -n 4 > $OUT/pytest_new.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_new.log
tail -40 $OUT/pytest_new.log | cut -c1-300
stamp new_tests
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -n 6 --deselect tests/test_zy_gpu_widening.py > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log | cut -c1-300
stamp old_tests
# 2. bench lines
B="timeout 300 python bench.py"
$B > $OUT/bench_potrf.json 2> $OUT/bench_potrf.err; tail -1 $OUT/bench_potrf.json | cut -c1-400
stamp bench_potrf
$B --routine posv_mixed --steps 2 > $OUT/bench_posv_mixed.json 2> $OUT/bench_posv_mixed.err; tail -1 $OUT/bench_posv_mixed.json | cut -c1-1200; tail -3 $OUT/bench_posv_mixed.err
stamp bench_posv_mixed
$B --routine gesv_mixed --steps 2 > $OUT/bench_gesv_mixed.json 2> $OUT/bench_gesv_mixed.err; tail -1 $OUT/bench_gesv_mixed.json | cut -c1-1200; tail -3 $OUT/bench_gesv_mixed.err
stamp bench_gesv_mixed
$B --routine tileops --steps 5 > $OUT/bench_tileops.json 2> $OUT/bench_tileops.err; tail -1 $OUT/bench_tileops.json | cut -c1-1500; tail -3 $OUT/bench_tileops.err
stamp bench_tileops
$B --routine getrf --no-cpu-baseline > $OUT/bench_getrf.json 2> $OUT/bench_getrf.err; tail -1 $OUT/bench_getrf.json | cut -c1-400
stamp bench_getrf
$B --routine zgemm --steps 2 > $OUT/bench_zgemm.json 2> $OUT/bench_zgemm.err; tail -1 $OUT/bench_zgemm.json | cut -c1-600; tail -3 $OUT/bench_zgemm.err
$B --routine zherk --steps 2 > $OUT/bench_zherk.json 2> $OUT/bench_zherk.err; tail -1 $OUT/bench_zherk.json | cut -c1-600; tail -3 $OUT/bench_zherk.err
stamp bench_z
$B --impl reference --steps 1 --warmup 0 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -1 $OUT/bench_reference.json | cut -c1-300
stamp bench_reference
# 3. smoke
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
stamp smoke
# 4. ncu: launch list of a mixed solve, full captures of the tcgen05 kernel and of the geadd tile kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_posv_mixed.csv \
    python scratch/prof_mixed.py 8192 > $OUT/ncu_launches.log 2>&1
stamp ncu_launches
timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 0 -c 3 -f -o $OUT/prof_tf32x3 \
    python scratch/prof_mixed.py 8192 > $OUT/ncu_tf32x3.log 2>&1
stamp ncu_tf32x3
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tile_foreach -s 1 -c 1 -f -o $OUT/prof_geadd \
    python scratch/prof_mixed.py 2048 > $OUT/ncu_geadd.log 2>&1
stamp ncu_geadd
# 5. if time is left: getrf panel variants and phases
timeout 240 python scratch/perf_probe.py 32768 512 > $OUT/perf_probe.log 2> $OUT/perf_probe.err; cat $OUT/perf_probe.log
stamp perf_probe
ls -la $OUT
