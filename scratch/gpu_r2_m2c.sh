#!/usr/bin/env bash
# 2-GPU: sb200_bcast_tiles + the whole parity check once more on the final library
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MGPU_SIZES="1000x128" MGPU_WIDEN=0 timeout 300 $TR --master-port 29521 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m2c_check.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m2c_check.log; grep -E "bcast_tiles|MGPU|exit|Error|error|FAIL" $OUT/r2m2c_check.log | tail -8 | cut -c1-250
