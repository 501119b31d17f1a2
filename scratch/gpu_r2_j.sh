#!/usr/bin/env bash
# Round 2 call J (N = 1): v4 LU base kernel with fused leaf updates + redux argmax, batched Frobenius / max norms,
# shim 'N','N' -> 'N','T' (distinct B tiles transposed once): parity, phases, drop-in testers, tile-op bandwidths
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2j_timeline.txt; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "getrf or gesv or lu or panel or permute or norm or reference_tester" > $OUT/r2j_pytest.log 2>&1; tail -3 $OUT/r2j_pytest.log; stamp pytest
for n in 16384 32768; do
  SB200_PHASES=1 SB200_VARIANTS=default timeout 300 python scratch/perf_variants.py getrf $n 512 2>> $OUT/r2j_perf_getrf.err | grep routine | cut -c1-200 | tee -a $OUT/r2j_perf_getrf.log
done; stamp perf_getrf
grep sb200_phases $OUT/r2j_perf_getrf.err | cut -c1-400
timeout 300 python scratch/bench_contended.py 512 > $OUT/r2j_contended.log 2> $OUT/r2j_contended.err; grep "LU panel" $OUT/r2j_contended.log; grep sb200_phases $OUT/r2j_contended.err | sed -n '1p;8p' | cut -c1-300; stamp contended
export OMP_NUM_THREADS=16 OPENBLAS_NUM_THREADS=1
for t in tester_cublas tester_sb200; do
  for r in gemm potrf getrf; do
    timeout 400 oracle/_ref/$t --target d --origin d --type d --dim 16384 --nb 512 --check n --ref n --repeat 3 $r > $OUT/r2j_${t}_$r.log 2>&1
    echo "$t $r: $(grep -E '^ +d' $OUT/r2j_${t}_$r.log | awk '{print $(NF-9), $(NF-8)}' | tr '\n' ' ')"; grep -E "^ +d" $OUT/r2j_${t}_$r.log | tail -2 | cut -c1-200
  done
done; stamp testers
timeout 300 oracle/_ref/tester_sb200 --target d --origin d --type d,z --dim 1000,2048 --nb 256 --check y --ref n gemm > $OUT/r2j_tester_sb200_gemm_check.log 2>&1; grep -E "pass|FAIL|failed" $OUT/r2j_tester_sb200_gemm_check.log | tail -5 | cut -c1-200; stamp tester_check
timeout 300 python bench.py --routine tileops --steps 5 --warmup 3 > $OUT/r2j_bench_tileops.json 2> $OUT/r2j_bench_tileops.err; python - <<'PYEOF'
import json
d = json.loads(open("gpurun_out/r2j_bench_tileops.json").read().strip().splitlines()[-1])
print({k: round(v["frac"], 3) for k, v in d["kernels"].items()})
PYEOF
stamp tileops
