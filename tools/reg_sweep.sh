#!/bin/bash
# tuning experiment: rebuild with different register caps and measure (run on the GPU box)
for R in ${REGS:-64 80 96 128}; do
  B2G_MAXREG=$R python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
  for B in ${BLOCKS:-256}; do
    echo -n "== maxrregcount $R step block $B: "
    B2G_STEP_BLOCK=$B python tools/prof_case.py 65536 2
  done
done
python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
