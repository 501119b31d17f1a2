#!/usr/bin/env bash
# host-side wall time of the driver calls at the metric's size on one GPU, by segment (workspace cache, parallel plan)
OUT=gpurun_out; mkdir -p $OUT
SB200_HOST_TIMES=1 timeout 600 python bench.py --size 65536 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-also > $OUT/r2n2_bench.json 2> $OUT/r2n2_bench.err
grep sb200_host_ms $OUT/r2n2_bench.err | tail -3; cut -c1-330 $OUT/r2n2_bench.json
SB200_GETRF_DIST=1 SB200_HOST_TIMES=1 timeout 600 python bench.py --routine getrf --size 32768 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > $OUT/r2n2_bench_getrf.json 2> $OUT/r2n2_bench_getrf.err
grep sb200_host_ms $OUT/r2n2_bench_getrf.err | tail -2; cut -c1-330 $OUT/r2n2_bench_getrf.json
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -n 4 -k "potrf or getrf or posv or gesv or trmm or solve or potrs or getrs" > $OUT/r2n2_pytest.log 2>&1; tail -3 $OUT/r2n2_pytest.log
