mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
timeout 600 python scripts/bench_configs.py C1 C2 C3 C5s C5m > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl
