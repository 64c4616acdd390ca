# after the last kernel change: parity tests, the C4 / C2 bench lines, the HITON capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_C4_n1.json 2> gpurun_out/bench_n1.err; cut -c1-200 gpurun_out/r02_bench_C4_n1.json; echo
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C2_n1.json 2>> gpurun_out/bench_n1.err; cut -c1-200 gpurun_out/r02_bench_C2_n1.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hiton_fz -s 2 -c 1 -o gpurun_out/prof_hiton_C4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-blocks 0 > gpurun_out/ncu_hiton_C4.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_hiton_C4.ncu-rep gpurun_out/r02_hiton_fz_C4
grep -E "Duration|Issue Slots Busy|Registers Per|No Eligible" gpurun_out/r02_hiton_fz_C4_details.txt | head -5
grep -E "dram__bytes_(read|write).sum," gpurun_out/r02_hiton_fz_C4_raw.csv
head -4 gpurun_out/r02_hiton_fz_C4_opmix.txt
