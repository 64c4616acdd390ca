mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/pytest_multi.log 2>&1
tail -3 gpurun_out/pytest_multi.log
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-1500
tail -3 gpurun_out/bench_n$N.err
