#!/usr/bin/env bash
# Round 2, second 8-GPU call: the default bench line (dpotrf + also dgetrf / dgemm + e2e, n = 65536) with the final
# kernels, then one-variable variants of the chain: broadcast as scatter + all-gather, lookahead depth 3, fewer NCCL channels.
#   gpurun --gpus 8 --timeout 900 -- 'bash scratch/gpu_r2_g8b.sh'
set -o pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2g8b_timeline.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
SB200_PHASES=1 timeout 600 $TR --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2g8b_bench_default.json 2> $OUT/r2g8b_bench_default.err
echo "bench default exit $?"; python - <<'PYEOF'
import json
d = json.loads(open("gpurun_out/r2g8b_bench_default.json").read().strip().splitlines()[-1])
print("dpotrf", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "dev", round(d["device_ms_per_step"], 1), "trail", round(d["roofline"]["trailing_ms_per_step"], 1), "panel", round(d["roofline"]["panel_stream_ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "check", d["check"]["pass"], "e2e", round(d["e2e"]["value"], 1))
for k, v in d["also"].items():
    print(k, round(v["value"], 1), "ms", round(v["ms_per_step"], 1), "dev", round(v.get("device_ms_per_step", 0), 1), "trail", round(v["roofline"]["trailing_ms_per_step"], 1), "panel", round(v["roofline"]["panel_stream_ms_per_step"], 1), "frac", round(v["roofline"]["frac"], 3), "check", v["check"]["pass"])
PYEOF
grep sb200_phases $OUT/r2g8b_bench_default.err | grep '"rank": 0' | tail -3 | cut -c1-400; stamp bench_default
port=29540
bench() {
  local tag=$1 r=$2; shift 2
  port=$((port + 1))
  SB200_PHASES=1 timeout 300 $TR --master-port $port bench.py --gpus 8 --routine $r --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also "$@" > $OUT/r2g8b_bench_${r}_$tag.json 2> $OUT/r2g8b_bench_${r}_$tag.err
  echo "bench $r $tag exit $?"; tail -1 $OUT/r2g8b_bench_${r}_$tag.json | cut -c1-330; grep sb200_phases $OUT/r2g8b_bench_${r}_$tag.err | grep '"rank": 0' | tail -1 | cut -c1-400
  stamp "bench $r $tag"
}
SB200_BCAST=1 bench sag potrf
SB200_BCAST=1 bench sag getrf
SB200_LOOKAHEAD=3 bench la3 potrf
NCCL_MAX_NCHANNELS=8 bench nch8 potrf
