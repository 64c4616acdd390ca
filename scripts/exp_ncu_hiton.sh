mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hiton_fz -c 1 -o gpurun_out/prof_hiton -f python scripts/perf_probe.py ${EXP_P:-10000} ${EXP_N:-2000} 24 > gpurun_out/ncu_hiton.log 2>&1
tail -3 gpurun_out/ncu_hiton.log
ls -la gpurun_out/
