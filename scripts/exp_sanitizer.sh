# compute-sanitizer passes over the round-2 kernels (small tests only; slow tool)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "tests/test_gpu_fznz.py::test_fznz_pairwise_subsets_hiton" "tests/test_gpu_fznz.py::test_fznz_subsets_gram_sizes" "tests/test_gpu_pairwise_sharded.py::test_sharded_pairwise_errors" "tests/test_gpu_fz.py::test_cor_matrix_tolerance" "tests/test_gpu_mi.py::test_discrete_test_subsets" -x -q -m gpu ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -8
( timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest "tests/test_gpu_fznz.py::test_fznz_pairwise_subsets_hiton" "tests/test_gpu_fznz.py::test_fznz_subsets_gram_sizes" "tests/test_gpu_fz.py::test_hiton_pc_matches_oracle" -x -q -m gpu ) > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/sanitizer_racecheck.log | head -12
