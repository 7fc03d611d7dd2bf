# CTA size of k_step (B2G_STEP_BLOCK): ms per 100 steps
for cfg in "cfg2 65536" "cfg3 65536" "cfg4 65536" "cfg2 262144"; do
  for b in 0 224 384 448 512; do
    if [ $b = 0 ]; then unset B2G_STEP_BLOCK; else export B2G_STEP_BLOCK=$b; fi
    echo -n "block $b: "; python tools/prof_cfg.py $cfg 3 2>&1 | grep -v Warn | cut -c1-62
  done
done
