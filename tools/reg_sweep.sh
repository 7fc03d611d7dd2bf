#!/bin/bash
# tuning experiment: rebuild with different register caps and measure (run on the GPU box)
for R in 64 80 96 128; do
  B2G_MAXREG=$R python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
  echo "== maxrregcount $R"
  python tools/sweep.py test_env.yaml 4096,65536 64,128 2>&1 | tail -4
done
python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
