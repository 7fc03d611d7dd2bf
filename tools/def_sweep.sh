#!/bin/bash
# tuning experiment: rebuild with extra -D flags and measure (run on the GPU box).  usage: def_sweep.sh "" "-DFOO" ...
for D in "$@"; do
  B2G_DEFINES="$D" python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
  echo "== defines [$D]"
  python tools/prof_case.py 65536 2; python tools/prof_cfg.py cfg4 65536 2; python tools/prof_cfg.py cfg3 65536 2
done
python -c "from box2d_sim_env_b200 import build; build.build(force=True)" > /dev/null 2>&1
