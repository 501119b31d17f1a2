#!/usr/bin/env bash
# N=1: full GPU parity suite + variant timing probe.  usage: gpurun --timeout 900 -- 'bash scratch/gpu_call1.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.csv 2>&1
timeout 500 python -m pytest tests -m gpu -x -q --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 240 python scratch/perf_probe.py 32768 512 > $OUT/probe.log 2> $OUT/probe.err; echo "probe exit $?" >> $OUT/probe.log
cat $OUT/probe.log; grep -E "^##|sb200_phases" $OUT/probe.err | cut -c1-900
