mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fz.py tests/test_gpu_fznz.py tests/test_gpu_properties.py -x -q -m gpu ) 2>&1 | tail -4
timeout 300 python scripts/perf_probe.py 50000 10000 24 2>&1 | grep -E "rep 1|rror" | sed -e 's/| pairwise.*//'
timeout 300 python scripts/bench_configs.py C5m 2>&1 | tail -1 | cut -c1-400
