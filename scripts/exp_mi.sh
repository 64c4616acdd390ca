mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mi.py tests/test_gpu_fznz.py tests/test_gpu_prep.py -x -q -m gpu ) 2>&1 | tail -6
timeout 300 python scripts/bench_configs.py C3 2>&1 | tail -1 | cut -c1-600
