#!/usr/bin/env bash
# round 2, last GPU minute: every test of the tournament / complex LU / complex mixed-solver work, the LU regressions
# around the host-code edits of getrf.cu, BASELINE configs[0] at full size; then one bench line each for zgetrf,
# getrf_tntpiv and dgetrf at n = 16384 on one GPU, and smoke()
OUT=gpurun_out; mkdir -p $OUT
timeout 50 python -u -m pytest -m gpu -q --tb=short --timeout 40 -n 8 -p no:cacheprovider \
  tests/test_zzzz_gpu_tntpiv.py tests/test_zzzz_gpu_complex_lu.py tests/test_gpu_drivers.py tests/test_zz_gpu_panel_variants.py \
  tests/test_zy_gpu_widening.py tests/test_zzz_gpu_round2_candidates.py tests/test_gpu_mixed.py \
  -k "(getrf or nopiv or gesv or getrs or lu_factor or tntpiv or complex or config0 or mixed) and not 2048 and not 4096 and not 1536 and not full_size_potrf and not base_kernel_variants_identical" \
  > $OUT/r2r2_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/r2r2_pytest.log | cut -c1-250
for r in zgetrf getrf_tntpiv getrf; do
  timeout 20 python bench.py --routine $r --n 16384 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > $OUT/r2r2_bench_$r.json 2> $OUT/r2r2_bench_$r.err
  echo "bench $r rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/r2r2_bench_$r.json").read().strip().splitlines()[-1])
    print("$r", round(d["value"], 2), d["unit"], "ms", round(d["ms_per_step"], 1), "check", d.get("check"))
except Exception as e:
    print("$r: no line", e)
PY
done
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/r2r2_smoke.log | cut -c1-300
