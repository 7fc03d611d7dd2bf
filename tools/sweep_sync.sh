# phase-barrier masks of k_step (B2G_SYNC_MASK, read when a batch is created): ms per 100 steps
for cfg in "cfg4 65536" "cfg2 65536" "cfg3 65536"; do
  for m in 0x0 0x30 0x2aa 0x3ff 0x7ff 0x1fff; do
    echo -n "mask $m: "; B2G_SYNC_MASK=$m python tools/prof_cfg.py $cfg 3 2>&1 | grep -v Warn | cut -c1-62
  done
done
