#!/usr/bin/env bash
# GPU parity suite only.  usage: gpurun --timeout 1500 -- 'bash scratch/gpu_tests.sh [pytest args]'
set -uo pipefail
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
