mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fznz.py tests/test_gpu_fz.py -x -q -m gpu ) > gpurun_out/exp_pytest.log 2>&1
tail -5 gpurun_out/exp_pytest.log
timeout 600 python scripts/bench_configs.py ${CONFIGS:-C5m} 2>&1 | tail -3
