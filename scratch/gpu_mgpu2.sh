#!/usr/bin/env bash
# 2-GPU sanity of the refactored multi-rank drivers (all types) + one short bench line per routine.
# usage: gpurun --gpus 2 --timeout 420 -- 'bash scratch/gpu_mgpu2.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MGPU_SIZES="1000x128,1024x256" timeout 240 $TR --master-port 29511 scratch/mgpu_check.py 1x2 2x1 > $OUT/mgpu2_check.log 2>&1
echo "mgpu_check exit $?" >> $OUT/mgpu2_check.log
grep -E "grid|MGPU|exit|Error|error" $OUT/mgpu2_check.log | tail -30
echo "[$((SECONDS-T0)) s] check"
timeout 150 $TR --master-port 29512 bench.py --gpus 2 --routine potrf --steps 2 --warmup 3 --no-e2e > $OUT/mgpu2_bench_potrf.json 2> $OUT/mgpu2_bench_potrf.err
echo "bench potrf exit $?"; tail -1 $OUT/mgpu2_bench_potrf.json | cut -c1-400; tail -3 $OUT/mgpu2_bench_potrf.err
echo "[$((SECONDS-T0)) s] potrf"
timeout 150 $TR --master-port 29513 bench.py --gpus 2 --routine zherk --size 32768 --steps 2 --no-e2e > $OUT/mgpu2_bench_zherk.json 2> $OUT/mgpu2_bench_zherk.err
echo "bench zherk exit $?"; tail -1 $OUT/mgpu2_bench_zherk.json | cut -c1-400; tail -3 $OUT/mgpu2_bench_zherk.err
echo "[$((SECONDS-T0)) s] zherk"
