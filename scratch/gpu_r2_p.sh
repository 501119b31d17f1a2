#!/usr/bin/env bash
# LU parity with the tightened factor bound (2e-12)
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q --timeout 200 -n 4 -k "getrf" > $OUT/r2p_pytest.log 2>&1; tail -4 $OUT/r2p_pytest.log | cut -c1-300
