#!/bin/bash
# Stand-in for compute-sanitizer (closed on this GPU pool): build the library with device-side bounds assertions
# (-DB2G_CHECK: every state / model / list access checked, violations print and trap) next to the product library.
# Then, on a GPU box:   B2G_LIB=$PWD/box2d_sim_env_b200/libb2g_check.so python tools/sanitize_case.py 6
#                       B2G_LIB=$PWD/box2d_sim_env_b200/libb2g_check.so python -m pytest tests -m gpu -x -q
set -e
cd "$(dirname "$0")/.."
B2G_DEFINES="-DB2G_CHECK" python -c "
from box2d_sim_env_b200 import build
print(build.build(force=True, out='$PWD/box2d_sim_env_b200/libb2g_check.so'))"
