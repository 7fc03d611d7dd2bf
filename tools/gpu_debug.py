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
    q, qd = random_states(batch.model, n, seed=21, spread=0.8)
    tg = random_targets(batch.model, 0, n, seed=22, segments=segs)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    check("set_state")
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
    print("status", np.unique(batch.status()))
    t0 = time.time(); batch.stepPhysics(100); print("100 steps x", n, "worlds:", time.time() - t0, "s", batch.take_kernel_ms())
print("ALL OK")
