# full GPU validation: parity tests, the bench line, and (optionally) the config table / ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
if [ -n "$CONFIGS" ]; then
  timeout 900 python scripts/bench_configs.py $CONFIGS > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
  cat gpurun_out/configs.jsonl
fi
if [ -n "$NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hiton_fz -s 2 -c 1 -o gpurun_out/prof_hiton_C4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_hiton_C4.log 2>&1
  timeout 900 ncu --set full --clock-control none -k regex:cor_tc2 -s 24 -c 1 -o gpurun_out/prof_cor_C4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_cor_C4.log 2>&1
  ls -la gpurun_out
fi
