#!/usr/bin/env bash
# 2-GPU bench lines of BASELINE config 3 (complex double gemmC / herk paths) at reduced n.
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29513 bench.py --gpus 2 --routine zherk --size 24576 --steps 2 --no-e2e > $OUT/mgpu3_bench_zherk.json 2> $OUT/mgpu3_bench_zherk.err
echo "bench zherk exit $?"; tail -1 $OUT/mgpu3_bench_zherk.json | cut -c1-500; tail -2 $OUT/mgpu3_bench_zherk.err
echo "[$((SECONDS-T0)) s] zherk"
timeout 100 $TR --master-port 29514 bench.py --gpus 2 --routine zgemm --size 16384 --steps 2 --no-e2e > $OUT/mgpu3_bench_zgemm.json 2> $OUT/mgpu3_bench_zgemm.err
echo "bench zgemm exit $?"; tail -1 $OUT/mgpu3_bench_zgemm.json | cut -c1-500; tail -2 $OUT/mgpu3_bench_zgemm.err
echo "[$((SECONDS-T0)) s] zgemm"
