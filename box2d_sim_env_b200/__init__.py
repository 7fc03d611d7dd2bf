"""box2d_sim_env_b200 — B200-native batched Box2D backend for the sim_env::World interface.

Host-side Python mirror of the reference's Box2DWorld API (include/sim_env/Box2DWorld.h:748-844 of
JoshuaHaustein/box2d_sim_env) over the C ABI in include/b2g.h.  All compute runs in hand-written CUDA
kernels (csrc/); there is no CPU fallback — creating a batch without a CUDA device raises.
"""
from .batch_world import (B2GError, BatchBox2DWorld, Environment, Model, MultiGpuBatch, data_path, lib, pinned_empty)  # noqa: F401
