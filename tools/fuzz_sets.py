"""Exploration: random per-tile parameter-set histories (b2g_batch_set_link_property on tile-aligned selections, whole-batch
setters in between, stepping, checkCollision) against oracle worlds that receive the same setter calls.
Usage: fuzz_sets.py <seed_lo> <seed_hi>   (env B2G_LANES / B2G_STEP_BLOCK choose the stepping kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import box2d_sim_env_b200 as b2g
import test_gpu_parity as t
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    env_name = [t.ENV_A, t.ENV_B][seed % 2]
    n = int(rng.integers(40, 200))
    batch, worlds = t.scattered_start(b2g, env_name, n, seed=1000 + seed)
    nb = batch.model.n_bodies
    links = [b for b in range(1, nb)]
    props = ["mass", "ground_friction", "ground_torque_integral", "contact_friction"]
    ntiles = (n + 31) // 32
    try:
        for it in range(int(rng.integers(4, 10))):
            body = int(rng.choice(links)); prop = str(rng.choice(props))
            value = float(rng.uniform(0.05, 2.0)) if prop != "ground_torque_integral" else float(rng.uniform(0.001, 0.1))
            if rng.uniform() < 0.25:
                first, count = 0, n
                batch.selectWorlds(0, 0)
            else:
                t0 = int(rng.integers(0, ntiles)); t1 = int(rng.integers(t0 + 1, ntiles + 1))
                first, count = 32 * t0, min(32 * t1, n) - 32 * t0
                batch.selectWorlds(first, count)
            batch.setLinkProperty(body, prop, value)
            batch.selectWorlds(0, 0)
            for ow in worlds[first:first + count]:
                ow.set_link_property(body, prop, value)
            steps = int(rng.integers(1, 12))
            batch.stepPhysics(steps)
            for ow in worlds:
                ow.step(steps)
            if rng.uniform() < 0.3:
                counts, cio, _ = batch.checkCollision(max_per_world=64)
                for w, ow in enumerate(worlds):
                    oio, _ = ow.check_collision()
                    assert counts[w] == len(oio) and np.array_equal(cio[w, :len(oio)], oio), "checkCollision differs, world %d" % w
            t.assert_same(batch, worlds, "seed %d after setter %d (%s body %d on [%d, %d))" % (seed, it, prop, body, first, first + count))
        w = int(rng.integers(0, n)); body = int(rng.choice(links)); prop = str(rng.choice(props))
        batch.selectWorlds(w, 1)
        assert batch.getLinkProperty(body, prop) == worlds[w].get_link_property(body, prop), "getter differs"
        batch.selectWorlds(0, 0)
    except AssertionError as e:
        bad += 1
        print("seed", seed, str(e)[:300])
print("parameter-set fuzz seeds %d..%d: %d failures" % (lo, hi, bad))
