#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the bench lines, the ncu launch list and one full capture of
# the dominant kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
# usage: gpurun --timeout 1500 -- 'bash scratch/gpu_round.sh [tests] [bench] [ncu]'
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
WHAT="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt

if [[ "$WHAT" == *tests* ]]; then
    timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
    echo "pytest exit $?" >> $OUT/pytest_gpu.log
    tail -5 $OUT/pytest_gpu.log
fi
if [[ "$WHAT" == *bench* ]]; then
    timeout 600 python bench.py > $OUT/bench_potrf.json 2> $OUT/bench_potrf.err
    tail -1 $OUT/bench_potrf.json
    timeout 600 python bench.py --routine getrf --no-cpu-baseline > $OUT/bench_getrf.json 2> $OUT/bench_getrf.err
    tail -1 $OUT/bench_getrf.json
    timeout 600 python bench.py --routine gemm --no-cpu-baseline > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err
    tail -1 $OUT/bench_gemm.json
    timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
    tail -1 $OUT/bench_reference.json
fi
if [[ "$WHAT" == *ncu* ]]; then
    # launch list of the bench command (cold-cache, serialised: compare SHARES)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_potrf.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1
    # full capture of the dominant kernel (trailing-update DMMA GEMM, nb = 512 tiles)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 2 -c 2 \
        -f -o $OUT/prof_gemm_dmma python scratch/prof_gemm.py > $OUT/ncu_full.log 2>&1
    ls -la $OUT
fi
