#!/usr/bin/env bash
# 4 GPUs (2 x 2, the grid of the N = 4 point of the scaling run): parity of every driver
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
SB200_BCAST_MIN=65536 MGPU_SIZES="1000x128" MGPU_WIDEN=0 timeout 200 $TR --master-port 29521 scratch/mgpu_check.py 2x2 > $OUT/r2m4_check.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m4_check.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2m4_check.log | tail -12 | cut -c1-250
