#!/bin/bash
# r2d: green-context SM reservation experiment + ncu captures of the GEMM variants actually timed
mkdir -p gpurun_out
T0=$(date +%s); tick() { echo "[$(( $(date +%s) - T0 )) s] $1" | tee -a gpurun_out/r2d_timeline.txt; }
timeout 300 python scratch/bench_greenctx.py 512 > gpurun_out/r2d_greenctx.log 2> gpurun_out/r2d_greenctx.err; tick greenctx
# potrf trailing variant <NT> inside a real dpotrf (n = 16384): full capture of two trailing launches
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 40 -c 2 -f -o gpurun_out/r2d_prof_potrf_gemm \
    python bench.py --routine potrf --size 16384 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also > gpurun_out/r2d_ncu_potrf.log 2>&1; tick ncu_potrf
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 200 -c 2 -f -o gpurun_out/r2d_prof_getrf_gemm \
    python bench.py --routine getrf --size 16384 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-also > gpurun_out/r2d_ncu_getrf.log 2>&1; tick ncu_getrf
