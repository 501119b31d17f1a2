#!/usr/bin/env bash
# FIRST GPU call of the next round: what the end of round 2 could not run.
#   gpurun --gpus 2 --timeout 900 -- 'bash scratch/gpu_r3_first.sh 2'        (then the same with --gpus 8 and the argument 8)
# 1. multi-rank parity with the tournament-pivoting LU on real grids (scratch/mgpu_check.py: getrf_tntpiv against the oracle,
#    which is pinned to the reference's own multi-rank runs, tests/golden/grid_getrf_tntpiv_d_*.npz)
# 2. the GPU tests added after the budget ended (the grid golden files against the GPU tournament) and the whole suite
# 3. bench lines: getrf_tntpiv on the grid, zgetrf on one GPU (both with the probe-vector LU check)
N=${1:-2}
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GRIDS=$([ "$N" = 8 ] && echo "2x4 4x2" || echo "1x$N ${N}x1")
MGPU_SIZES="1000x128,1024x256,2048x256" timeout 500 $TR --master-port 29621 scratch/mgpu_check.py $GRIDS > $OUT/r3a_check_${N}gpu.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r3a_check_${N}gpu.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r3a_check_${N}gpu.log | tail -60 | cut -c1-250
timeout 200 python -m pytest -m gpu -q --timeout 100 -n 4 tests/test_zzzz_gpu_tntpiv.py tests/test_zzzz_gpu_complex_lu.py > $OUT/r3a_pytest_new.log 2>&1; tail -3 $OUT/r3a_pytest_new.log
timeout 300 $TR --master-port 29622 bench.py --gpus $N --routine getrf_tntpiv --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r3a_bench_getrf_tntpiv_${N}gpu.json 2> $OUT/r3a_bench_getrf_tntpiv_${N}gpu.err
tail -1 $OUT/r3a_bench_getrf_tntpiv_${N}gpu.json | cut -c1-400
timeout 120 python bench.py --routine zgetrf --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r3a_bench_zgetrf_1gpu.json 2> $OUT/r3a_bench_zgetrf_1gpu.err
tail -1 $OUT/r3a_bench_zgetrf_1gpu.json | cut -c1-400
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -n 4 > $OUT/r3a_pytest_all.log 2>&1; tail -3 $OUT/r3a_pytest_all.log
