#!/usr/bin/env bash
# tile kernels after the load-all-then-store skeleton: parity + bandwidths
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_reference_tester_gpu.py -m gpu -q -x --timeout 120 -n 4 -k "geadd or gescale or geset or gecopy or trapezoid or single_tile or row_col or unit or tester" > $OUT/r2o_pytest.log 2>&1; tail -2 $OUT/r2o_pytest.log
timeout 200 python bench.py --routine tileops --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2o_bench_tileops.json 2> $OUT/r2o_bench_tileops.err; python - <<'PYEOF'
import json
d = json.loads(open("gpurun_out/r2o_bench_tileops.json").read().strip().splitlines()[-1])
print({k: round(v["frac"], 3) for k, v in d["kernels"].items()})
PYEOF
