#!/usr/bin/env bash
# Round 2, 8-GPU call (8x the box time: keep it short).  2 x 4 parity of every driver, the default bench line (dpotrf +
# also dgetrf / dgemm + e2e, n = 65536) with phase timers, BASELINE configs[3] (zgemm / zherk n = 40960) and configs[4]
# (dgesv_mixed / dposv_mixed n = 65536), dpotrf without the chain SM partition for comparison.
#   gpurun --gpus 8 --timeout 900 -- 'bash scratch/gpu_r2_g8.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2g8_timeline.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MGPU_WIDEN=0 MGPU_SIZES="2048x256" timeout 300 $TR --master-port 29531 scratch/mgpu_check.py 2x4 > $OUT/r2g8_check.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2g8_check.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2g8_check.log | tail -12 | cut -c1-250; stamp check
SB200_PHASES=1 timeout 600 $TR --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2g8_bench_default.json 2> $OUT/r2g8_bench_default.err
echo "bench default exit $?"; tail -1 $OUT/r2g8_bench_default.json | cut -c1-3000; grep sb200_phases $OUT/r2g8_bench_default.err | grep '"rank": 0' | tail -4 | cut -c1-400; stamp bench_default
port=29540
bench() {
  local tag=$1 r=$2; shift 2
  port=$((port + 1))
  SB200_PHASES=1 timeout 300 $TR --master-port $port bench.py --gpus 8 --routine $r --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also "$@" > $OUT/r2g8_bench_${r}_$tag.json 2> $OUT/r2g8_bench_${r}_$tag.err
  echo "bench $r $tag exit $?"; tail -1 $OUT/r2g8_bench_${r}_$tag.json | cut -c1-500; grep sb200_phases $OUT/r2g8_bench_${r}_$tag.err | grep '"rank": 0' | tail -1 | cut -c1-400
  stamp "bench $r $tag"
}
bench default gesv_mixed --size 65536
bench default posv_mixed --size 65536
bench default zgemm --size 40960
bench default zherk --size 40960
SB200_CHAIN_SMS=0 bench nochain potrf
