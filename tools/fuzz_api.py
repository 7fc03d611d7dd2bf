"""Exploration: random sequences of API calls (step, targets, efforts, DOF subsets, poses, link enable/disable, link
mass / ground friction, collision checks) on a batch, mirrored on oracle worlds, full state compared after every call.  Usage: fuzz_api.py <first_seed> <last_seed>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import box2d_sim_env_b200 as b2g
import test_gpu_parity as t
from parity_utils import ENV_A, ENV_B, ENV_2, ENV_3, env_dict, oracle_worlds, random_states


def run(seed, env_name, n=12, n_ops=40):
    rng = np.random.default_rng(seed)
    batch = t.make_batch(b2g, env_name, n, max_contacts=64)
    worlds = oracle_worlds(env_dict(env_name), n)
    model = batch.model
    lim = model.dof_limits()
    q, qd = random_states(model, n, seed=seed, spread=0.7)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    log = []
    for k in range(n_ops):
        op = int(rng.integers(0, 12))
        if op == 0:
            s = int(rng.integers(1, 12)); sleep = bool(rng.integers(0, 2))
            batch.stepPhysics(s, True, sleep)
            for ow in worlds: ow.step(s, True, sleep)
            log.append("step %d sleep=%s" % (s, sleep))
        elif op == 1:
            nd = model.objects[0]["n_dofs"]
            tg = rng.uniform(lim[:nd, 2] * 1.5, lim[:nd, 3] * 1.5, (n, nd)).astype(np.float32)   # partly beyond the limits: scaled
            batch.setTargetVelocity(0, tg)
            for w, ow in enumerate(worlds): ow.set_target_velocity(0, tg[w])
            log.append("target")
        elif op == 2:
            nd = model.objects[0]["n_dofs"]
            e = rng.uniform(-2, 2, (n, nd)).astype(np.float32); e[:, 3:] *= 0.05
            batch.setControlEfforts(0, e)
            for w, ow in enumerate(worlds): ow.set_control_efforts(0, e[w])
            log.append("efforts")
        elif op in (3, 4):
            o = int(rng.integers(0, model.n_objects)); info = model.objects[o]; nd = info["n_dofs"]
            if nd == 0: continue
            kk = int(rng.integers(1, nd + 1)); idx = np.sort(rng.choice(nd, kk, replace=False)).astype(np.int32)
            lo = np.maximum(lim[info["dof_offset"] + idx, 0], -1.5); hi = np.minimum(lim[info["dof_offset"] + idx, 1], 1.5)
            vals = rng.uniform(lo, hi, (n, kk)).astype(np.float32)
            vel = op == 4
            if vel: vals = rng.uniform(-3, 3, (n, kk)).astype(np.float32)
            (batch.setDOFVelocities if vel else batch.setDOFPositions)(o, vals, idx)
            for w, ow in enumerate(worlds): ow.set_object_dofs(o, vals[w], idx, velocities=vel)
            log.append("dofs obj %d idx %s vel=%s" % (o, idx.tolist(), vel))
        elif op == 5:
            o = int(rng.integers(0, model.n_objects))
            pose = np.concatenate([rng.uniform(-1, 1, (n, 2)), rng.uniform(-3, 3, (n, 1))], 1).astype(np.float32)
            batch.setObjectPose(o, pose)
            for w, ow in enumerate(worlds): ow.set_object_pose(o, *[float(x) for x in pose[w]])
            log.append("pose obj %d" % o)
        elif op == 6:
            body = int(rng.integers(1, model.n_bodies)); en = rng.integers(0, 2, n).astype(np.uint8)
            batch.setLinkEnabled(body, en)
            for w, ow in enumerate(worlds): ow.set_link_enabled(body, bool(en[w]))
            log.append("enable body %d" % body)
        elif op == 11:
            body = int(rng.integers(1, model.n_bodies))
            prop = ["mass", "ground_friction", "ground_torque_integral"][int(rng.integers(0, 3))]
            value = float(np.float32(rng.uniform(0.02, 1.5) if prop == "mass" else rng.uniform(0.0, 0.6)))
            batch.setLinkProperty(body, prop, value)
            for ow in worlds: ow.set_link_property(body, prop, value)
            log.append("link %d %s = %g" % (body, prop, value))
        elif op == 8:
            s = int(rng.integers(1, 10))
            counts, io, fo = batch.stepPhysicsRecord(s, max_per_world=64)
            for w, ow in enumerate(worlds):
                oio, ofo = ow.step_record(s)
                assert counts[w] == len(oio) and np.array_equal(io[w, :len(oio)], oio) and np.array_equal(fo[w, :len(oio)], ofo), \
                    "recorded contacts differ, world %d after %s" % (w, log[-6:])
            log.append("stepRecord %d" % s)
        elif op == 9:
            box = np.sort(rng.uniform(-1.5, 1.5, (2, 2)), axis=0).astype(np.float32)
            aabb = [float(box[0, 0]), float(box[0, 1]), float(box[1, 0]), float(box[1, 1])]
            excl = bool(rng.integers(0, 2))
            hits = batch.getObjects(aabb, exclude_robots=excl)
            for w, ow in enumerate(worlds):
                assert np.array_equal(hits[w], ow.objects_in_aabb(aabb, excl, model.n_objects)), "getObjects(aabb) differs, world %d" % w
            log.append("getObjects")
        elif op == 10:
            feas = batch.isPhysicallyFeasible()
            for w, ow in enumerate(worlds):
                assert bool(feas[w]) == bool(ow.is_physically_feasible()), "isPhysicallyFeasible differs, world %d" % w
            log.append("feasible")
        else:
            counts, io, fo = batch.checkCollision(max_per_world=64)
            for w, ow in enumerate(worlds):
                oio, ofo = ow.check_collision()
                assert counts[w] == len(oio) and np.array_equal(io[w, :len(oio)], oio) and np.array_equal(fo[w, :len(oio)], ofo), \
                    "checkCollision differs, world %d after %s" % (w, log[-6:])
            log.append("checkCollision")
        t.assert_same(batch, worlds, "seed %d %s after %s" % (seed, env_name, log[-6:]))
    assert np.all((batch.status() & 3) == 0), "status %s" % batch.status()   # bit 2 only notes a target beyond the limits


if __name__ != "__main__":
    raise SystemExit
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(lo, hi):
    env_name = [ENV_A, ENV_B, ENV_2, ENV_3][seed % 4]
    try:
        run(seed, env_name)
    except b2g.B2GError as e:
        print('seed', seed, 'B2GError:', str(e)[:120])
    except AssertionError as e:
        bad.append((seed, env_name, str(e)[:600]))
print("API fuzz seeds %d..%d: %d failures" % (lo, hi, len(bad)))
for b in bad[:4]:
    print(b)
