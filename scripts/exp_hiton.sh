# kernel-tuning experiment: parity tests of the fz path, then the hiton phase of C2 / C4 with alternative builds
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fz.py tests/test_gpu_properties.py -x -q -m gpu ) > gpurun_out/exp_pytest.log 2>&1
tail -5 gpurun_out/exp_pytest.log
for so in "" $(ls build/exp/*.so 2>/dev/null); do
  echo "== lib: ${so:-default}"
  FW_LIB_PATH=${so:+$PWD/$so} timeout 300 python scripts/perf_probe.py ${EXP_P:-50000} ${EXP_N:-10000} 24 2>&1 | grep -E "rep 1|Error|error"
done
