#!/usr/bin/env bash
# Round 2 call K (N = 1): the whole GPU suite after the clean-up, then the default bench line
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
T0=$SECONDS; stamp() { echo "[$((SECONDS-T0)) s] $*" | tee -a $OUT/r2k_timeline.txt; }
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -n 4 > $OUT/r2k_pytest.log 2>&1; tail -15 $OUT/r2k_pytest.log; stamp pytest
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; stamp smoke
