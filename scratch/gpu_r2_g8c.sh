#!/usr/bin/env bash
# Round 2, third 8-GPU call: the default bench line with scatter + all-gather broadcasts and NCCL capped at 8 CTAs
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
SB200_PHASES=1 timeout 400 $TR --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2g8c_bench_default.json 2> $OUT/r2g8c_bench_default.err
echo "bench default exit $?"; python - <<'PYEOF'
import json
d = json.loads(open("gpurun_out/r2g8c_bench_default.json").read().strip().splitlines()[-1])
print("dpotrf", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "dev", round(d["device_ms_per_step"], 1), "trail", round(d["roofline"]["trailing_ms_per_step"], 1), "panel", round(d["roofline"]["panel_stream_ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "check", d["check"]["pass"], "e2e", round(d["e2e"]["value"], 1))
for k, v in d["also"].items():
    print(k, round(v["value"], 1), "ms", round(v["ms_per_step"], 1), "dev", round(v.get("device_ms_per_step", 0), 1), "trail", round(v["roofline"]["trailing_ms_per_step"], 1), "panel", round(v["roofline"]["panel_stream_ms_per_step"], 1), "frac", round(v["roofline"]["frac"], 3), "check", v["check"]["pass"])
PYEOF
tail -3 $OUT/r2g8c_bench_default.err | cut -c1-300
