# compute-sanitizer passes over the new kernels (small tests only; slow tool)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "tests/test_gpu_fznz.py::test_fznz_pairwise_subsets_hiton" "tests/test_gpu_fznz.py::test_fznz_subsets_gram_sizes" "tests/test_gpu_prep.py::test_reference_fixtures" "tests/test_gpu_mi.py::test_discrete_test_subsets" -x -q -m gpu ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -8
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest "tests/test_gpu_fz.py::test_hiton_pc_large_accepted_sets" -x -q -m gpu -k "26" ) > gpurun_out/sanitizer_racecheck.log 2>&1
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest "tests/test_gpu_fz.py::test_hiton_pc_matches_oracle" "tests/test_gpu_fznz.py::test_fznz_pairwise_subsets_hiton" "tests/test_gpu_mi.py::test_discrete_test_subsets" -x -q -m gpu ) > gpurun_out/sanitizer_racecheck2.log 2>&1
echo "racecheck2 rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported" gpurun_out/sanitizer_racecheck2.log | head -12
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | head -8
