#!/usr/bin/env bash
# 2-GPU parity after the communicator changes: scatter + all-gather broadcasts forced on small ranges too, capped
# communicators for the Cholesky chain
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SB200_BCAST_MIN=65536 MGPU_SIZES="1000x128,2048x256" timeout 400 $TR --master-port 29521 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m2b_check.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m2b_check.log; grep -E "MGPU|exit|Error|error|FAIL" $OUT/r2m2b_check.log | tail -8 | cut -c1-250; grep -c " ok" $OUT/r2m2b_check.log
