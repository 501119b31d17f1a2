#!/usr/bin/env bash
# Multi-GPU parity + bench lines at N ranks.  usage: gpurun --gpus N --timeout 900 -- 'bash scratch/gpu_mgpu.sh N [check] [bench]'
set -o pipefail
MGPU_SHAPES="${MGPU_SHAPES:-}"
N="${1:-2}"; shift || true
WHAT="${*:-check bench}"
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_n$N.csv 2>&1
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1
if [[ "$WHAT" == *check* ]]; then
    timeout 420 $TR --master-port 29511 scratch/mgpu_check.py $MGPU_SHAPES > $OUT/mgpu_check_n$N.log 2>&1
    echo "mgpu_check exit $?" >> $OUT/mgpu_check_n$N.log
    grep -E "grid|MGPU|exit|Error|error" $OUT/mgpu_check_n$N.log | tail -30
fi
if [[ "$WHAT" == *bench* ]]; then
    for r in potrf getrf gemm; do
        timeout 420 $TR --master-port 29512 bench.py --gpus $N --routine $r --steps 2 --warmup 3 > $OUT/bench_${r}_n$N.json 2> $OUT/bench_${r}_n$N.err
        echo "bench $r exit $?"; tail -1 $OUT/bench_${r}_n$N.json | cut -c1-600; tail -3 $OUT/bench_${r}_n$N.err
    done
fi
