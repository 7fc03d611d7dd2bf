"""Throughput sweep over batch size / block size on one GPU (device-timed rollout kernel only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import box2d_sim_env_b200 as b2g
import bench

env = sys.argv[1] if len(sys.argv) > 1 else "test_env.yaml"
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4096,16384,65536,262144").split(",")]
blocks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "32,64,128").split(",")]
model = b2g.Model(b2g.data_path(env))
limits = model.dof_limits()
for n in sizes:
    for blk in blocks:
        os.environ["B2G_BLOCK"] = str(blk)
        batch = b2g.BatchBox2DWorld(n).loadWorld(model)
        if env == "test_env.yaml":
            q, qd, tg = bench.synth_inputs(limits, n, 1)
        else:
            rng = np.random.default_rng(1)
            q, qd = batch.getWorldState()
            nd = model.objects[0]["n_dofs"]
            tg = rng.uniform(limits[:nd, 2], limits[:nd, 3], (10, n, nd)).astype(np.float32)
        best = 1e30
        for rep in range(3):
            batch.setWorldState(q, qd, reset_caches=True)
            batch.take_kernel_ms()
            for t in range(10):
                batch.setTargetVelocity(0, tg[t])
                batch.stepPhysics(10)
            ms, nl = batch.take_kernel_ms()
            best = min(best, ms)
        print("%s n=%d block=%d: %.3f ms / 100 steps -> %.3e world-steps/s" % (env, n, blk, best, n * 100 / (best * 1e-3)), flush=True)
        del batch
