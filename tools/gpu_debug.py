"""GPU-box debugging aid: step oracle and CUDA side by side and print the first differences."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import box2d_sim_env_b200 as b2g
from parity_utils import *

continuous = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
segs = int(sys.argv[3]) if len(sys.argv) > 3 else 6
for env_name in (ENV_A, ENV_B):
    print("=====", env_name, "continuous", continuous)
    batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.data_path(env_name))
    worlds = oracle_worlds(load_env_dict(env_name), n, continuous=bool(continuous))

    def check(what):
        bodies = batch.bodies(); contacts = batch.contacts(); fat = batch.fat_aabbs(); jio, jfo = batch.joints()
        nbad = 0
        for w, ow in enumerate(worlds):
            d = compare_world(bodies, contacts, w, ow)
            if not np.array_equal(fat[w], ow.fat_aabbs()): d.append("fat differ")
            oji, ojf = ow.joints()
            if not np.array_equal(jio[w], oji): d.append("joint ints differ")
            if not np.array_equal(jfo[w][:, :6], ojf[:, :6]): d.append("joint floats differ %s" % np.argwhere(jfo[w][:, :6] != ojf[:, :6]).tolist()[:4])
            if d:
                nbad += 1
                if nbad <= 3:
                    print("  world", w, d[:8])
        print(what, "bad worlds:", nbad, "/", n, "contacts:", int(contacts[0].sum()), "touching:", int(contacts[1][..., 2].sum()))
        return nbad

    check("fresh")
    q, qd = random_states(batch.model, n, seed=21, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=22, segments=segs)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    check("set_state")
    # put the static obstacle near the origin so that dynamic-vs-static contacts (TOI candidates) happen
    rng = np.random.default_rng(5)
    poses = np.concatenate([rng.uniform(-0.6, 0.6, (n, 2)), rng.uniform(-3.1, 3.1, (n, 1))], axis=1).astype(np.float32)
    batch.setObjectPose("obj2", poses)
    for w, ow in enumerate(worlds):
        ow.set_object_pose(2, *[float(x) for x in poses[w]])
    check("set obj2 pose")
    for seg in range(segs):
        batch.setTargetVelocity(0, tg[seg])
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
        for s in range(10):
            batch.stepPhysics(1)
            for ow in worlds:
                ow.step(1)
            if check("seg %d step %d" % (seg, s)):
                sys.exit(1)
        print("   toi events this segment (oracle):", sum(ow.stats()["toi_events"] for ow in worlds), "calls", sum(ow.stats()["toi_calls"] for ow in worlds))
    print("status", np.unique(batch.status()))
    t0 = time.time(); batch.stepPhysics(100); print("100 steps x", n, "worlds:", time.time() - t0, "s", batch.take_kernel_ms())
print("ALL OK")
