#!/usr/bin/env bash
# round 2, last GPU minutes, call 1: one test of every new kind (tournament pivoting, complex LU / solve) + the LU regressions
# most exposed to the host-code edits of getrf.cu (from_real scalars, PanelScratch layout, BaseArgs in the shared header)
OUT=gpurun_out; mkdir -p $OUT
timeout 48 python -u -m pytest -m gpu -v --tb=short --timeout 30 -n 4 -p no:cacheprovider \
  "tests/test_gpu_drivers.py::test_getrf_matches_reference_golden_with_identical_pivots" \
  "tests/test_zz_gpu_panel_variants.py::test_getrf_panel_variants_identical_pivots[700-300-128-1-0]" \
  "tests/test_zz_gpu_panel_variants.py::test_getrf_panel_variants_identical_pivots[700-300-128-2-0]" \
  "tests/test_zzz_gpu_round2_candidates.py::test_getrf_nopiv_zero_pivot_info_and_pivoting_is_back_afterwards" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_matches_reference_golden_with_identical_pivots" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_tournament_matches_oracle[384-384-64-2]" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_tournament_matches_oracle[300-300-64-3]" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_tournament_matches_oracle[448-256-64-4]" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_grid_algorithm_on_one_rank[300-300-64-2]" \
  "tests/test_zzzz_gpu_tntpiv.py::test_tntpiv_rejects_the_shapes_the_reference_rejects" \
  "tests/test_zzzz_gpu_tntpiv.py::test_lu_factor_method_dispatch" \
  "tests/test_zzzz_gpu_complex_lu.py::test_zgetrf_matches_reference_golden_with_identical_pivots" \
  "tests/test_zzzz_gpu_complex_lu.py::test_cgetrf_against_reference_golden" \
  "tests/test_zzzz_gpu_complex_lu.py::test_zgetrf_vs_oracle_and_tester_residual[700-128]" \
  "tests/test_zzzz_gpu_complex_lu.py::test_zgesv_matches_reference_golden" \
  "tests/test_zzzz_gpu_complex_lu.py::test_zgetrf_rectangular_and_zero_pivot_info" \
  > $OUT/r2r1_pytest.log 2>&1
echo "rc=$?"; grep -E "PASSED|FAILED|ERROR|passed|failed" $OUT/r2r1_pytest.log | cut -c1-200 | tail -40
