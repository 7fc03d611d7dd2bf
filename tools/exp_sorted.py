"""Go / no-go experiment for activity binning: time one rollout launch with the worlds in input order and again with the
worlds sorted by how contact-rich they turn out to be (a-posteriori key from the first run), so that the warps are
homogeneous.  Usage: exp_sorted.py <config> <worlds>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import box2d_sim_env_b200 as b2g
import bench
name, n = sys.argv[1], int(sys.argv[2])
cfg = bench.CONFIGS[name]
model = bench.load_model(cfg)
batch = b2g.BatchBox2DWorld(n).loadWorld(model)
bench.place_obstacle(batch, cfg, n)
q, qd, tg = bench.synth_inputs(model.dof_limits(), n, 1, cfg)

def run(q, qd, tg, reps=3):
    d_tg = torch.from_numpy(np.ascontiguousarray(tg)).cuda()
    for r in range(reps):
        batch.setWorldState(q, qd, reset_caches=True)
        batch.take_kernel_ms()
        batch.rollout_dev(0, d_tg.data_ptr(), bench.N_TARGETS, bench.STEPS_PER_TARGET)
        ms, nl = batch.take_kernel_ms()
    counts, cio, _ = batch.contacts(max_per_world=80)
    return ms / nl, counts, cio[..., 2].sum(axis=1)

ms0, live, touching = run(q, qd, tg)
print("%s n=%d input order: %.3f ms; live %.2f touching %.2f; worlds with live %.3f with touching %.3f" %
      (name, n, ms0, live.mean(), touching.mean(), (live > 0).mean(), (touching > 0).mean()))
# mid-rollout key: contacts after 50 steps would be better; use the end-of-rollout key plus the initial "resting" flag
key = touching * 100 + live
order = np.argsort(-key, kind="stable")
ms1, _, _ = run(q[order], qd[order], tg[:, order])
print("sorted by end-of-rollout (touching, live): %.3f ms  (x%.2f)" % (ms1, ms0 / ms1))
