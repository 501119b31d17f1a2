#!/usr/bin/env bash
# Round 2, 8-GPU call (after the 1-GPU and 2-GPU scripts are green; 8x the box time, so keep it short):
#   gpurun --gpus 8 --timeout 900 -- 'bash scratch/gpu_r2_8gpu.sh [SWITCHES...]'
# e.g.  bash scratch/gpu_r2_8gpu.sh SB200_DIAG_MW=1 SB200_TRSM_FUSED=7 SB200_PANEL_LL=1
# Runs the 2x4 parity check and the headline bench lines (n = 65536) with the DEFAULT paths and, if switches are given,
# again with them; then the BASELINE configs[3] / [4] lines that have no 8-GPU number yet.
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
SW="$*"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29530
check() {  # tag [env...]
  local tag=$1; shift
  port=$((port + 1))
  env "$@" MGPU_WIDEN=0 MGPU_SIZES="2048x256" timeout 300 $TR --master-port $port scratch/mgpu_check.py 2x4 > $OUT/r2g8_check_$tag.log 2>&1
  echo "mgpu_check $tag exit $?" | tee -a $OUT/r2g8_check_$tag.log; grep -E "grid|MGPU|Error|error|FAIL" $OUT/r2g8_check_$tag.log | tail -20
  echo "[$((SECONDS-T0)) s] check $tag"
}
bench() {  # tag routine [env...]
  local tag=$1 r=$2; shift 2
  port=$((port + 1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 8 --routine $r --steps 2 --warmup 3 > $OUT/r2g8_bench_${r}_$tag.json 2> $OUT/r2g8_bench_${r}_$tag.err
  echo "bench $r $tag exit $?"; tail -1 $OUT/r2g8_bench_${r}_$tag.json | cut -c1-360; tail -2 $OUT/r2g8_bench_${r}_$tag.err
  echo "[$((SECONDS-T0)) s] bench $r $tag"
}
check default
for r in potrf getrf gemm; do bench default $r; done
if [ -n "$SW" ]; then
  check candidates $SW
  for r in potrf getrf gemm; do bench candidates $r $SW; done
fi
for r in zgemm zherk; do bench default $r; done                      # BASELINE configs[3], n = 40960
if [ "${SB200_RUN_UNVALIDATED:-0}" = "1" ]; then                      # configs[4]: needs the p x q solve path (gpu_r2_mgpu.sh green)
  for r in gesv_mixed posv_mixed; do bench default $r SB200_RUN_UNVALIDATED=1; done
fi
