"""Small fixed case for ncu: one rollout launch of `steps` physics steps on n worlds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import box2d_sim_env_b200 as b2g
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
model = b2g.Model(b2g.data_path("test_env.yaml"))
batch = b2g.BatchBox2DWorld(n).loadWorld(model)
q, qd, tg = bench.synth_inputs(model.dof_limits(), n, 1)
for r in range(reps):
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, tg[0])
    batch.stepPhysics(steps)
ms, nl = batch.take_kernel_ms()
print("n=%d steps=%d: %.3f ms per launch -> %.3e world-steps/s" % (n, steps, ms / nl, n * steps * nl / (ms * 1e-3)))
