#!/usr/bin/env bash
# Round 2, 2-GPU call: multi-rank parity of every driver (incl. the p x q solve path, the widening routines, the chain
# SM partition next to NCCL, the v4 LU base kernel), then bench lines with phase timers.
#   gpurun --gpus 2 --timeout 900 -- 'bash scratch/gpu_r2_m2.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2m2_timeline.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MGPU_SIZES="1000x128,1024x256,2048x256" timeout 400 $TR --master-port 29521 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m2_check.log 2>&1
rc=$?; echo "mgpu_check exit $rc" >> $OUT/r2m2_check.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2m2_check.log | tail -40 | cut -c1-250; stamp check
if [ $rc -ne 0 ]; then
  SB200_CHAIN_SMS=0 MGPU_SIZES="1000x128,1024x256" timeout 400 $TR --master-port 29522 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m2_check_nochain.log 2>&1
  echo "mgpu_check (no chain partition) exit $?" >> $OUT/r2m2_check_nochain.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2m2_check_nochain.log | tail -40 | cut -c1-250; stamp check_nochain
fi
port=29530
bench() {  # tag routine [extra args...] -- env via leading VAR=val before the call
  local tag=$1 r=$2; shift 2
  port=$((port + 1))
  SB200_PHASES=1 timeout 300 $TR --master-port $port bench.py --gpus 2 --routine $r --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also "$@" > $OUT/r2m2_bench_${r}_$tag.json 2> $OUT/r2m2_bench_${r}_$tag.err
  echo "bench $r $tag exit $?"; tail -1 $OUT/r2m2_bench_${r}_$tag.json | cut -c1-420; grep sb200_phases $OUT/r2m2_bench_${r}_$tag.err | tail -2 | cut -c1-400
  stamp "bench $r $tag"
}
bench default potrf
SB200_CHAIN_SMS=0 bench nochain potrf
bench default getrf
bench default gemm
SB200_RUN_UNVALIDATED=1 bench default posv_mixed --size 32768
SB200_RUN_UNVALIDATED=1 bench default gesv_mixed --size 32768
bench default zherk --size 24576
