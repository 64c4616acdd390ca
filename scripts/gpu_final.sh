# round-end evidence on ONE GPU: parity tests, bench lines of every configuration, the reference arm, launch list and ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_C4_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/r02_bench_C4_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_C4_reference_arm.json 2>> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/r02_bench_C4_reference_arm.json
for c in C2 C3 C5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02_bench_${c}_n1.json 2>> gpurun_out/bench_n1.err; cut -c1-220 gpurun_out/r02_bench_${c}_n1.json; echo; done
timeout 600 python scripts/bench_configs.py C1 C2 C3 C5 > gpurun_out/r02_configs.jsonl 2> gpurun_out/configs.err; cut -c1-500 gpurun_out/r02_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_C4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-blocks 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hiton_fz -s 2 -c 1 -o gpurun_out/prof_hiton_C4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-blocks 0 > gpurun_out/ncu_hiton_C4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cor_tc3 -s 24 -c 1 -o gpurun_out/prof_cor3_C4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-blocks 0 > gpurun_out/ncu_cor_C4.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_hiton_C4.ncu-rep gpurun_out/r02_hiton_fz_C4
bash scripts/ncu_export.sh gpurun_out/prof_cor3_C4.ncu-rep gpurun_out/r02_cor_tc3_C4
python - <<'PY'
import csv, json
def dram(path):
    rd = wr = None
    for r in csv.reader(open(path)):
        if len(r) >= 4 and r[1] == "dram__bytes_read.sum": rd = float(r[3]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[2]]
        if len(r) >= 4 and r[1] == "dram__bytes_write.sum": wr = float(r[3]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[2]]
    return rd, wr
out = {"source": "ncu --set full --clock-control none captures of bench.py --steps 1 --warmup 1 (scripts/gpu_final.sh): dram__bytes_read.sum + dram__bytes_write.sum per launch"}
rd, wr = dram("gpurun_out/r02_hiton_fz_C4_raw.csv"); out["hiton_fz_kernel_C4_dram_bytes_per_launch"] = rd + wr; out["hiton_fz_kernel_C4_read_write"] = [rd, wr]
rd, wr = dram("gpurun_out/r02_cor_tc3_C4_raw.csv"); out["cor_tc3_kernel_C4_dram_bytes_per_launch"] = rd + wr; out["cor_tc3_kernel_C4_read_write"] = [rd, wr]
json.dump(out, open("gpurun_out/r02_traffic.json", "w"), indent=1); print(out)
PY
grep -E "Duration|Issue Slots Busy|Registers Per" gpurun_out/r02_hiton_fz_C4_details.txt | head -4
grep -E "Duration|TC is|L2 Hit" gpurun_out/r02_cor_tc3_C4_details.txt | head -4
ls -la gpurun_out | tail -30
