#!/usr/bin/env bash
# N=1: tcgen05 parity tests + per-phase timing of the potrf / getrf critical path (SB200_PHASES=1).
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_mixed.py -m gpu -x -q > $OUT/pytest_mixed.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_mixed.log
tail -15 $OUT/pytest_mixed.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
for n in 8192 32768; do
  SB200_PHASES=1 timeout 200 $B --routine potrf --n $n > $OUT/ph_potrf_$n.json 2> $OUT/ph_potrf_$n.err; grep sb200_phases $OUT/ph_potrf_$n.err | tail -1; cut -c1-200 $OUT/ph_potrf_$n.json
  SB200_PHASES=1 timeout 200 $B --routine getrf --n $n > $OUT/ph_getrf_$n.json 2> $OUT/ph_getrf_$n.err; grep sb200_phases $OUT/ph_getrf_$n.err | tail -1; cut -c1-200 $OUT/ph_getrf_$n.json
  SB200_GETRF_DIST=1 SB200_PHASES=1 timeout 200 $B --routine getrf --n $n > $OUT/ph_getrfdist_$n.json 2> $OUT/ph_getrfdist_$n.err; grep sb200_phases $OUT/ph_getrfdist_$n.err | tail -1; cut -c1-200 $OUT/ph_getrfdist_$n.json
done
