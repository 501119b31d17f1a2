#!/usr/bin/env bash
# Round 2, 2-GPU call (after scratch/gpu_r2_call1.sh is green on one GPU): multi-rank parity of every driver with the
# default paths (incl. the p x q solve path written at the end of round 1), then with every round-2 candidate switched on,
# then short bench lines for both.      gpurun --gpus 2 --timeout 900 -- 'bash scratch/gpu_r2_mgpu.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ALL="SB200_DIAG_MW=1 SB200_TILE_FUSED=1 SB200_TRSM_FUSED=7 SB200_PANEL_LL=1 SB200_GEMM_BT=1 SB200_PANEL_SKINNY=1"
MGPU_WIDEN=0 MGPU_SIZES="1000x128,1024x256" timeout 300 $TR --master-port 29521 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m_check_default.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m_check_default.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2m_check_default.log | tail -30
echo "[$((SECONDS-T0)) s] check default"
env $ALL MGPU_WIDEN=0 MGPU_SIZES="1000x128,1024x256" timeout 300 $TR --master-port 29522 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m_check_candidates.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m_check_candidates.log; grep -E "grid|MGPU|exit|Error|error|FAIL" $OUT/r2m_check_candidates.log | tail -30
echo "[$((SECONDS-T0)) s] check candidates"
# SURVEY 8(f) items 2-3 on the grid (her2k, syrk, syr2k, getrf_nopiv), separately so that a failure there cannot mask the above
MGPU_SOLVE=0 MGPU_SIZES="1000x128" timeout 300 $TR --master-port 29525 scratch/mgpu_check.py 1x2 2x1 > $OUT/r2m_check_widen.log 2>&1
echo "mgpu_check exit $?" >> $OUT/r2m_check_widen.log; grep -E "her2k|exit|Error|error|FAIL" $OUT/r2m_check_widen.log | tail -10
echo "[$((SECONDS-T0)) s] check widen"
for r in potrf getrf gemm; do
  timeout 200 $TR --master-port 29523 bench.py --gpus 2 --routine $r --steps 2 --warmup 3 --no-e2e > $OUT/r2m_bench_${r}_default.json 2> $OUT/r2m_bench_${r}_default.err
  echo "bench $r default exit $?"; tail -1 $OUT/r2m_bench_${r}_default.json | cut -c1-300
  env $ALL timeout 200 $TR --master-port 29524 bench.py --gpus 2 --routine $r --steps 2 --warmup 3 --no-e2e > $OUT/r2m_bench_${r}_candidates.json 2> $OUT/r2m_bench_${r}_candidates.err
  echo "bench $r candidates exit $?"; tail -1 $OUT/r2m_bench_${r}_candidates.json | cut -c1-300
  echo "[$((SECONDS-T0)) s] bench $r"
done
