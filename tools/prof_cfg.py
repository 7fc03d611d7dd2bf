"""ncu case for any bench config: one bench step (set_state + one k_step rollout launch) on n worlds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import box2d_sim_env_b200 as b2g
import bench
name = sys.argv[1]
n = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = bench.CONFIGS[name]
model = bench.load_model(cfg)
batch = b2g.BatchBox2DWorld(n).loadWorld(model)
bench.place_obstacle(batch, cfg, n)
q, qd, tg = bench.synth_inputs(model.dof_limits(), n, 1, cfg)
d_tg = torch.from_numpy(tg).cuda()
for r in range(reps):
    batch.setWorldState(q, qd, reset_caches=True)
    batch.take_kernel_ms()
    batch.rollout_dev(0, d_tg.data_ptr(), bench.N_TARGETS, bench.STEPS_PER_TARGET)
    ms, nl = batch.take_kernel_ms()
counts, cio, _ = batch.contacts(max_per_world=80)
print("%s n=%d: %.3f ms per 100-step rollout launch -> %.3e world-steps/s; live contacts/world %.2f touching %.2f" %
      (name, n, ms / nl, n * 100 * nl / (ms * 1e-3), counts.mean(), cio[..., 2].sum() / n))
