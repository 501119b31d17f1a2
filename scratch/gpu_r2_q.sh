#!/usr/bin/env bash
# mixed-precision solvers on one GPU with the final LU panel / Cholesky chain (BASELINE configs[4] shape at n = 32768)
OUT=gpurun_out; mkdir -p $OUT
for r in gesv_mixed posv_mixed; do
  timeout 200 python bench.py --routine $r --size 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2q_bench_$r.json 2> $OUT/r2q_bench_$r.err
  python - <<PYEOF
import json
d = json.loads(open("gpurun_out/r2q_bench_$r.json").read().strip().splitlines()[-1])
print("$r", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), d.get("phases_ms"), d.get("iterations"), d["roofline"]["frac"])
PYEOF
done
