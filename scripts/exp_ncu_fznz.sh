mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fznz_prefilter -c 1 -o gpurun_out/prof_fznz_prefilter -f python scripts/bench_configs.py C5m > gpurun_out/ncu_fznz1.log 2>&1
tail -2 gpurun_out/ncu_fznz1.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hiton_fz_kernel -c 1 -o gpurun_out/prof_hiton_fznz -f python scripts/bench_configs.py C5s > gpurun_out/ncu_fznz2.log 2>&1
tail -2 gpurun_out/ncu_fznz2.log | cut -c1-300
( timeout 600 python -m pytest tests/test_gpu_prep.py -x -q -m gpu ) 2>&1 | tail -3
timeout 900 python scripts/bench_configs.py PREP 2>&1 | tail -6
ls -la gpurun_out | head -20
