#!/usr/bin/env bash
# Round 2 call E (N = 1): one-round LU base kernel (getrf_base_v3.cu, default on) and the chain SM partition
# (sm_partition.cu, forced on one rank with SB200_CHAIN_SMS): parity first, then phases, then ncu of the timed GEMMs.
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2e_timeline.txt; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "getrf or gesv or lu or panel or permute or nopiv" > $OUT/r2e_pytest_lu.log 2>&1; tail -3 $OUT/r2e_pytest_lu.log; stamp pytest_lu
SB200_CHAIN_SMS=4 timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "potrf or posv" > $OUT/r2e_pytest_chain4.log 2>&1; tail -3 $OUT/r2e_pytest_chain4.log; stamp pytest_chain4
for v3 in 1 0; do for n in 16384 32768; do
  SB200_PANEL_V3=$v3 SB200_PHASES=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py getrf $n 512 2>> $OUT/r2e_perf_getrf.err | grep routine | sed "s/^/V3=$v3 /" | cut -c1-200 | tee -a $OUT/r2e_perf_getrf.log
done; done; stamp perf_getrf
grep sb200_phases $OUT/r2e_perf_getrf.err | cut -c1-400
for c in 0 4 8; do for n in 8192 16384 32768; do
  SB200_CHAIN_SMS=$c SB200_PHASES=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py potrf $n 512 2>> $OUT/r2e_perf_potrf.err | grep routine | sed "s/^/CHAIN=$c /" | cut -c1-200 | tee -a $OUT/r2e_perf_potrf.log
done; done; stamp perf_potrf
grep sb200_phases $OUT/r2e_perf_potrf.err | cut -c1-400
timeout 300 python scratch/bench_contended.py 512 > $OUT/r2e_contended.log 2> $OUT/r2e_contended.err; grep "LU panel" $OUT/r2e_contended.log; stamp contended
# ncu: the trailing-update launch of one step in the middle of a real factorisation / SUMMA (cudaProfilerStart window)
for r in potrf getrf gemm; do
  SB200_NCU_TIMED=12 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 6 -f -o $OUT/r2e_prof_${r}_trailing \
    python bench.py --routine $r --size 16384 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-also > $OUT/r2e_ncu_$r.log 2>&1; stamp ncu_$r
done
ls -la $OUT/*.ncu-rep
