# HITON phase at C4 (bench.py's device-timed region) for the default library and build/exp/*.so (excluding debug builds)
for so in "" $(ls build/exp/*.so 2>/dev/null | grep -v dbg); do
  echo "== lib: ${so:-default}"
  FW_LIB_PATH=${so:+$PWD/$so} timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-blocks 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms_per_step %.2f  kernel_ms %.2f  e2e_ms %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
