#!/bin/bash
# tuning: CTA size of the stepping kernel vs batch size (run on the GPU box)
for N in ${SIZES:-4096 16384 65536 262144}; do
  for B in ${BLOCKS:-32 64 128 256 512 1024}; do
    echo -n "step block $B: "
    B2G_STEP_BLOCK=$B python tools/prof_cfg.py ${CFG:-cfg2} $N 2
  done
done
