"""One rollout launch on round-1 style inputs of test_env.yaml (box 1.3 m ahead of the robot: contact-free), independent of
bench.py, so that libraries of different commits can be compared (run from the root of the tree to measure):
  B2G_LANES=7 python tools/prof_r1inputs.py [worlds]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import box2d_sim_env_b200 as b2g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = b2g.Model(b2g.data_path("test_env.yaml"))
batch = b2g.BatchBox2DWorld(n).loadWorld(model)
lim = np.asarray(model.dof_limits(), np.float32)[:7, 2:4]
rng = np.random.Generator(np.random.Philox(key=1))
q = np.zeros((n, model.nq), np.float32); qd = np.zeros_like(q)
q[:, 7] = rng.uniform(-0.17, 0.17, n) - 0.05; q[:, 8] = 0.2167 + 1.3
tg = rng.uniform(lim[:, 0], lim[:, 1], (10, n, 7)).astype(np.float32)
d = torch.from_numpy(tg).cuda()
for r in range(4):
    batch.setWorldState(q, qd, reset_caches=True); batch.take_kernel_ms()
    batch.rollout_dev(0, d.data_ptr(), 10, 10); ms, nl = batch.take_kernel_ms()
print("%s n=%d: %.3f ms per 100-step rollout launch" % (os.path.basename(os.getcwd()), n, ms / nl))
