# A/B timing of fw_cor_matrix at C4 between the default library and build/exp/*.so (alternating, several rounds: the kernel is power-limited)
for round in 1 2 3; do
  for so in "" $(ls build/exp/*.so 2>/dev/null | grep -v dbg); do
    echo "== round $round lib: ${so:-default}"
    FW_LIB_PATH=${so:+$PWD/$so} timeout 200 python scripts/cor3_check.py worker 2 50000 10000 /tmp/x.npy 2>&1 | grep "^mode"
  done
done
