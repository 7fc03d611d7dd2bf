"""CUDA path vs CPU oracle on identical inputs, through the C ABI (ctypes).  The arithmetic is IEEE single
precision in the same operation order on both sides (--fmad=false, shared deterministic sin/cos), so the bar is
BIT-EXACT for poses, velocities, fat AABBs, joint impulses, contact sets, manifold ids and impulses — far inside
the 1e-5 single-step tolerance BASELINE.json states."""
import os

import numpy as np
import pytest

from parity_utils import (ENV_2, ENV_3, ENV_A, ENV_B, base_front_edge, compare_world, env_dict, feasible_states, load_env_dict, oracle_worlds,
                          random_states, random_targets)

pytestmark = pytest.mark.gpu


def make_batch(b2g, env_name, n, max_contacts=0):
    src = b2g.Model(env_dict(env_name)) if env_name.startswith("variant:") else b2g.data_path(env_name)
    return b2g.BatchBox2DWorld(n, max_contacts=max_contacts).loadWorld(src)


def assert_same(batch, worlds, what, atol=0.0):
    bodies = batch.bodies()
    contacts = batch.contacts(max_per_world=128)
    fat = batch.fat_aabbs()
    jio, jfo = batch.joints()
    bad = []
    for w, ow in enumerate(worlds):
        d = compare_world(bodies, contacts, w, ow, atol)
        if not np.array_equal(fat[w], ow.fat_aabbs()):
            d.append("fat AABBs differ")
        oji, ojf = ow.joints()
        if not np.array_equal(jio[w], oji):
            d.append("joint ints differ: %s vs %s" % (jio[w].tolist(), oji.tolist()))
        if not np.array_equal(jfo[w][:, :6], ojf[:, :6]):
            d.append("joint impulses / motors differ")
        if d:
            bad.append((w, d[:6]))
    assert not bad, "%s: %d/%d worlds differ, first: %s" % (what, len(bad), len(worlds), bad[:2])


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3])
def test_fresh_world_state(b2g, env_name):
    """loadWorld leaves the same bodies, fat AABBs and DOF state as the oracle's createWorld."""
    batch = make_batch(b2g, env_name, 3)
    worlds = oracle_worlds(env_dict(env_name), 3)
    assert_same(batch, worlds, "fresh world")
    q, qd = batch.getWorldState()
    for w, ow in enumerate(worlds):
        oq, oqd = ow.get_state()
        assert np.array_equal(q[w], oq) and np.array_equal(qd[w], oqd)


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3])
def test_set_state_parity(b2g, env_name):
    n = 64
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    q, qd = random_states(batch.model, n, seed=7)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    assert_same(batch, worlds, "set_state")
    gq, gqd = batch.getWorldState()
    for w, ow in enumerate(worlds):
        oq, oqd = ow.get_state()
        assert np.array_equal(gq[w], oq) and np.array_equal(gqd[w], oqd)


def scatter_obstacle(batch, worlds, seed, spread=0.6):
    """Move the static obstacle obj2 near the origin (per-world random pose) so that dynamic-vs-static contacts,
    i.e. TOI candidates, occur."""
    n = len(worlds)
    rng = np.random.default_rng(seed)
    poses = np.concatenate([rng.uniform(-spread, spread, (n, 2)), rng.uniform(-3.1, 3.1, (n, 1))], axis=1).astype(np.float32)
    batch.setObjectPose("obj2", poses)
    for w, ow in enumerate(worlds):
        ow.set_object_pose(2, *[float(x) for x in poses[w]])


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3])
def test_single_step_parity(b2g, env_name):
    """fresh world -> set_state -> one stepPhysics with the velocity controller running (BASELINE.json's
    single-step gate: <= 1e-5 on poses and velocities, bit-exact contact sets; here everything is bit-exact)."""
    n = 256
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    q, qd = random_states(batch.model, n, seed=11, spread=0.6)
    tgt = random_targets(batch.model, 0, n, seed=12)[0]
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, tgt)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_target_velocity(0, tgt[w])
    scatter_obstacle(batch, worlds, seed=13)
    batch.stepPhysics(1)
    for ow in worlds:
        ow.step(1)
    assert_same(batch, worlds, "single step")
    gq, gqd = batch.getWorldState()
    oq = np.stack([ow.get_state()[0] for ow in worlds])
    oqd = np.stack([ow.get_state()[1] for ow in worlds])
    assert np.max(np.abs(gq - oq)) <= 1e-5 and np.max(np.abs(gqd - oqd)) <= 1e-5
    counts, cio, _ = batch.contacts()
    assert int(cio[..., 2].sum()) > 20, "scenario produced too few touching contacts"


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3])
def test_multi_step_trajectory_parity(b2g, env_name):
    """120 steps with the target resampled every 10 steps: contact creation order, warm starting, sleeping and
    TOI sub-steps all have to agree for the trajectories to stay bit-identical."""
    n = 128
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    q, qd = random_states(batch.model, n, seed=21, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=22, segments=12)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    scatter_obstacle(batch, worlds, seed=23)
    touching_seen = 0
    toi_events = 0
    for seg in range(12):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "segment %d" % seg)
        counts, cio, _ = batch.contacts()
        touching_seen += int(cio[..., 2].sum())
        toi_events += sum(ow.stats()["toi_events"] for ow in worlds)
    assert touching_seen > 100, "scenario produced too few touching contacts; parity would be vacuous"
    if not env_name.startswith("variant:"):   # env2 / env3 have no static body: continuous collision never triggers
        assert toi_events > 0, "scenario never exercised the TOI sub-step path"
    assert np.all(batch.status() == 0)


def test_check_collision_parity(b2g):
    n = 256
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=31, spread=0.6)
    batch.setWorldState(q, qd, reset_caches=True)
    counts, io, fo = batch.checkCollision(max_per_world=32)
    total = 0
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        oio, ofo = ow.check_collision()
        assert counts[w] == len(oio), "world %d: %d vs %d touching contacts" % (w, counts[w], len(oio))
        assert np.array_equal(io[w, :len(oio)], oio)
        assert np.array_equal(fo[w, :len(oio)], ofo)
        total += len(oio)
    assert total > 20
    assert_same(batch, worlds, "after checkCollision")


def test_sleeping_and_no_sleep_flag(b2g):
    n = 8
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=41, vel=0.0)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    batch.stepPhysics(70, execute_controllers=False)
    for ow in worlds:
        ow.step(70, execute_controllers=False)
    assert_same(batch, worlds, "sleep")
    batch.stepPhysics(3, execute_controllers=False, allow_sleeping=False)
    for ow in worlds:
        ow.step(3, execute_controllers=False, allow_sleeping=False)
    assert_same(batch, worlds, "allow_sleeping=False")


def test_save_restore_and_set_pose(b2g):
    n = 16
    batch = make_batch(b2g, ENV_A, n)
    q0, qd0 = batch.getWorldState()
    batch.saveState()
    batch.setTargetVelocity("my_robot", [0.5, 0.0, 0.3, 0.5, -0.5, -0.5, 0.5])
    batch.stepPhysics(20)
    q1, _ = batch.getWorldState()
    assert not np.array_equal(q0, q1)
    assert batch.restoreState() is True
    q2, qd2 = batch.getWorldState()
    assert np.allclose(q2, q0, atol=1e-6) and np.allclose(qd2, qd0, atol=1e-5)
    assert batch.restoreState() is False          # empty stack: the reference returns false (:3526-3528)
    worlds = oracle_worlds(load_env_dict(ENV_A), 1)
    batch.setObjectPose("obj2", [2.0, -1.0, 0.7])
    worlds[0].set_object_pose(2, 2.0, -1.0, 0.7)
    assert np.array_equal(batch.getObjectPose("obj2")[0], worlds[0].get_object_pose(2))


def test_determinism_across_launches(b2g):
    n = 512
    a = make_batch(b2g, ENV_A, n)
    b = make_batch(b2g, ENV_A, n)
    q, qd = random_states(a.model, n, seed=51, spread=0.7)
    tg = random_targets(a.model, 0, n, seed=52)[0]
    for batch in (a, b):
        batch.setWorldState(q, qd, reset_caches=True)
        batch.setTargetVelocity(0, tg)
    a.stepPhysics(60)
    for _ in range(6):
        b.stepPhysics(10)
    assert np.array_equal(a.bodies(), b.bodies())


def test_error_behaviour(b2g):
    batch = make_batch(b2g, ENV_A, 4)
    with pytest.raises(b2g.B2GError):
        batch.setWorldState(np.zeros((4, 3)), np.zeros((4, 3)))      # reference: std::runtime_error (:1700-1704)
    with pytest.raises(b2g.B2GError):
        batch.setTargetVelocity("obj1", np.zeros((4, 3)))            # only robots have a velocity controller


@pytest.mark.parametrize("lpw", [None, "1", "8"])
def test_clutter_shape_c_parity(b2g, lpw, monkeypatch):
    """Shape C (BASELINE.json configs[3]): robot + 32 convex polygons, 38 bodies / 39 proxies / 41 joints per world.
    Dense sliding clutter: many simultaneous contacts, multi-body islands, contact creation/destruction order.  A small batch
    of a scene this size steps with a group of lanes per world by default (b2g_capi.cpp); one thread per world and another
    group size are forced through B2G_LPW."""
    from box2d_sim_env_b200 import envgen
    if lpw is not None:
        monkeypatch.setenv("B2G_LPW", lpw)
    n = 24
    env = envgen.clutter_env()
    batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.Model(env))
    worlds = oracle_worlds(env, n)
    assert_same(batch, worlds, "fresh clutter world")
    q, qd = envgen.clutter_states(n, seed=5, speed=3.0)
    tg = random_targets(batch.model, 0, n, seed=6, segments=6)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    touching = 0
    for seg in range(6):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "clutter segment %d" % seg)
        counts, cio, _ = batch.contacts(max_per_world=80)
        touching += int(cio[..., 2].sum())
    assert touching > 30, "clutter scenario produced too few touching contacts"
    assert np.all(batch.status() == 0)
    # batched checkCollision on the same worlds
    counts, io, fo = batch.checkCollision(max_per_world=80)
    for w, ow in enumerate(worlds):
        oio, ofo = ow.check_collision()
        assert counts[w] == len(oio) and np.array_equal(io[w, :len(oio)], oio) and np.array_equal(fo[w, :len(oio)], ofo)


def test_maximum_size_world_parity(b2g):
    """A world at the capacity limit (63 bodies, 64 proxies, 66 joints: robot + 57 clutter objects), narrow-warp and
    full-warp kernels: bit-exact vs the oracle over a contact-rich rollout."""
    from box2d_sim_env_b200 import envgen
    env = envgen.clutter_env(n_objects=57)
    model = b2g.Model(env)
    assert model.n_fixtures == 64
    n = 12
    batch = b2g.BatchBox2DWorld(n, max_contacts=128).loadWorld(model)
    worlds = oracle_worlds(env, n)
    # robot at the origin, objects on a 0.4 m grid around it (none closer than 0.7 m), sliding fast at random
    rng = np.random.default_rng(9)
    cells = [(x, y) for x in np.arange(-1.8, 1.81, 0.4) for y in np.arange(-1.8, 1.81, 0.4) if x * x + y * y > 0.7 ** 2][:57]
    assert len(cells) == 57
    q = np.zeros((n, model.nq), np.float32)
    qd = np.zeros_like(q)
    for k, (x, y) in enumerate(cells):
        q[:, 7 + 3 * k] = x + rng.uniform(-0.05, 0.05, n)
        q[:, 8 + 3 * k] = y + rng.uniform(-0.05, 0.05, n)
        q[:, 9 + 3 * k] = rng.uniform(-3, 3, n)
        qd[:, 7 + 3 * k:9 + 3 * k] = rng.uniform(-3, 3, (n, 2))
        qd[:, 9 + 3 * k] = rng.uniform(-3, 3, n)
    tg = random_targets(model, 0, n, seed=10, segments=3)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    for seg in range(3):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "max-size world, segment %d" % seg)
    assert np.all(batch.status() == 0)
    counts, cio, _ = batch.contacts(max_per_world=128)
    assert int(cio[..., 2].sum()) > 20


@pytest.mark.parametrize("n", [1, 31, 33, 95])
def test_ragged_batch_sizes(b2g, n):
    """Batch sizes that do not fill the last 32-world tile: the padding lanes must never be touched or read."""
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=61, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=62)[0]
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, tg)
    batch.stepPhysics(15)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_target_velocity(0, tg[w])
        ow.step(15)
    assert_same(batch, worlds, "ragged batch n=%d" % n)


def scattered_start(b2g, env_name, n, seed):
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    q, qd = random_states(batch.model, n, seed=seed, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=seed + 1)[0]
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, tg)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_target_velocity(0, tg[w])
    return batch, worlds


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B])
def test_recorded_contact_stepping(b2g, env_name):
    """stepPhysics(std::vector<Contact>&, steps) (reference :3290-3310, startRecordings :2743-2768): the per-world record of
    BeginContact events — pairs, order, contact points and normals — is identical to the oracle's."""
    n = 96
    batch, worlds = scattered_start(b2g, env_name, n, seed=71)
    counts, io, fo = batch.stepPhysicsRecord(25, max_per_world=32)
    total = 0
    for w, ow in enumerate(worlds):
        oio, ofo = ow.step_record(25)
        assert counts[w] == len(oio), "world %d: %d vs %d recorded contacts" % (w, counts[w], len(oio))
        assert np.array_equal(io[w, :len(oio)], oio) and np.array_equal(fo[w, :len(oio)], ofo)
        total += len(oio)
    assert total > 30, "scenario recorded too few contacts"
    assert_same(batch, worlds, "after recorded stepping")


def test_guarded_stepping(b2g):
    """stepPhysics(callback, steps) (reference :3312-3333) with the callback as a body-pair table."""
    n, steps = 128, 40
    batch, worlds = scattered_start(b2g, ENV_A, n, seed=81)
    nb = batch.model.n_bodies
    ok, done = batch.stepPhysicsGuarded(np.ones((nb, nb), np.uint8), steps)
    assert ok.all() and (done == steps).all()
    for ow in worlds:
        ow.step(steps)
    assert_same(batch, worlds, "guarded stepping with every pair allowed")
    # forbid every pair: a world stops after the step in which its first BeginContact fires
    batch, worlds = scattered_start(b2g, ENV_A, n, seed=83)
    ok, done = batch.stepPhysicsGuarded(np.zeros((nb, nb), np.uint8), steps)
    assert (~ok).sum() > 10 and ok.sum() > 0, "scenario must abort some worlds and complete others"
    assert np.all(done[ok] == steps) and np.all(done[~ok] >= 1)
    touching_at_abort = 0
    for w, ow in enumerate(worlds):
        ow.step(int(done[w]))
        oio, _ = ow.contacts()
        if not ok[w] and len(oio) and oio[:, 2].any():
            touching_at_abort += 1
        if ok[w]:
            assert not (len(oio) and oio[:, 2].any()), "world %d completed although a contact is touching" % w
    assert_same(batch, worlds, "guarded stepping: state after the executed steps")
    assert touching_at_abort >= 0.8 * (~ok).sum()


def test_status_bits(b2g):
    """Per-world status: bit0 = contact-slot capacity exceeded (loud: B2G_ERR_CAPACITY), bit1 = non-finite state."""
    from box2d_sim_env_b200 import envgen
    n = 8
    env = envgen.clutter_env()
    batch = b2g.BatchBox2DWorld(n, max_contacts=2).loadWorld(b2g.Model(env))   # far too few slots for sliding clutter
    q, qd = envgen.clutter_states(n, seed=5, speed=3.0)
    batch.setWorldState(q, qd, reset_caches=True)
    batch.stepPhysics(30)
    with pytest.raises(b2g.B2GError) as err:
        batch.status()
    assert err.value.code == 5 and "capacity" in str(err.value)
    batch = make_batch(b2g, ENV_A, 4)
    q, qd = batch.getWorldState()
    q[2, 0] = np.nan          # base x of the robot in world 2
    batch.setWorldState(q, qd)
    batch.stepPhysics(2)
    assert batch.status().tolist() == [0, 0, 2, 0]


def test_link_enable_disable_rest(b2g):
    """Box2DLink::setEnabled (reference :802-822): a disabled link stops colliding (contacts are filtered out on the next
    Collide), re-enabling brings the pairs back; plus setToRest / atRest."""
    n = 64
    batch, worlds = scattered_start(b2g, ENV_A, n, seed=91)
    batch.stepPhysics(10)
    for ow in worlds:
        ow.step(10)
    obj1_body = batch.model.objects[1]["first_body"]
    flags = (np.arange(n) % 2 == 0)
    batch.setLinkEnabled(obj1_body, ~flags)          # disable obj1 in the even worlds
    for w, ow in enumerate(worlds):
        ow.set_link_enabled(obj1_body, not flags[w])
    batch.stepPhysics(15)
    for ow in worlds:
        ow.step(15)
    assert_same(batch, worlds, "after disabling obj1 in half of the worlds")
    counts, cio, _ = batch.contacts()
    fix_of_obj1 = {f for f, row in enumerate(batch.model.fixture_model()[0]) if row[0] == obj1_body}
    for w in np.nonzero(flags)[0]:
        assert not any(int(f) in fix_of_obj1 for f in cio[w, :counts[w], :2].reshape(-1)), "disabled link still has contacts"
    batch.setLinkEnabled(obj1_body, True)
    for ow in worlds:
        ow.set_link_enabled(obj1_body, True)
    batch.stepPhysics(15)
    for ow in worlds:
        ow.step(15)
    assert_same(batch, worlds, "after re-enabling obj1")
    assert not batch.atRest().all()
    batch.setToRest()
    assert batch.atRest().all()
    gq, gqd = batch.getWorldState()
    assert not gqd.any()


def test_region_query_and_feasibility(b2g):
    """getObjects(aabb) (reference :3207-3254) and isPhysicallyFeasible (:3370-3400) against the oracle on worlds whose
    objects are scattered over (and into) each other."""
    n = 192
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=95, spread=0.5)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    seen = set()
    for box, excl in (([-0.3, -0.3, 0.3, 0.3], False), ([0.0, 0.0, 1.0, 1.0], True), ([-3.0, -3.0, 3.0, 3.0], False),
                      ([5.0, 5.0, 6.0, 6.0], False)):
        hits = batch.getObjects(box, exclude_robots=excl)
        for w, ow in enumerate(worlds):
            assert np.array_equal(hits[w], ow.objects_in_aabb(box, exclude_robots=excl)), "world %d box %s" % (w, box)
        seen |= {tuple(r) for r in hits.tolist()}
    assert len(seen) >= 4, "region queries should see several different hit patterns"
    feas = batch.isPhysicallyFeasible()
    ofeas = np.array([ow.is_physically_feasible() for ow in worlds])
    assert np.array_equal(feas, ofeas)
    assert 0 < feas.sum() < n, "scenario must contain feasible and infeasible worlds"
    assert_same(batch, worlds, "after isPhysicallyFeasible (runs UpdateContacts)")


@pytest.mark.parametrize("env_name,n,step_block", [(ENV_A, 20000, None), (ENV_B, 12345, None), (ENV_A, 19999, 224), (ENV_B, 20001, 96),
                                                   (ENV_A, 95001, None), (ENV_B, 96000, 896)])
def test_large_batch_full_warp_path(b2g, env_name, n, step_block, monkeypatch):
    """Batches large enough for the full-warp, CTA-synchronised stepping kernel (phase barriers, parked padding worlds in
    the last CTA; small batches use the narrow-warp kernel), with the CTA size of the heuristic and with explicit
    non-power-of-two CTA sizes: final states of a 60-step rollout, all worlds, bit-exact.  The two batches beyond 94720
    worlds run the second instantiation of the kernel (72 registers, CTAs of up to 896 threads)."""
    from oracle import b2o
    if step_block:
        monkeypatch.setenv("B2G_STEP_BLOCK", str(step_block))
    batch = make_batch(b2g, env_name, n)
    nd = batch.model.objects[0]["n_dofs"]
    q, qd = random_states(batch.model, n, seed=97, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=98, segments=6)
    batch.setWorldState(q, qd, reset_caches=True)
    for seg in range(6):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
    gq, gqd = batch.getWorldState()
    ob = b2o.OracleBatch(env_dict(env_name), n)
    oq, oqd = ob.rollout(q, qd, 0, tg, 10, os.cpu_count() or 4)
    bad = np.nonzero(~(np.all(gq == oq, axis=1) & np.all(gqd == oqd, axis=1)))[0]
    assert bad.size == 0, "%d of %d worlds differ, first %s" % (bad.size, n, bad[:5].tolist())
    counts, cio, _ = batch.contacts(max_per_world=4)
    assert int(cio[..., 2].sum()) > 50
    assert np.all(batch.status() == 0)


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3])
def test_object_dof_subsets_parity(b2g, env_name):
    """Box2DObject::setDOFPositions / setDOFVelocities / getDOF*(values, indices) on index subsets (:1585-1626, 1691-1715):
    every object, several ascending subsets, interleaved with stepping; bit-exact vs the oracle."""
    n = 32
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    model = batch.model
    rng = np.random.default_rng(31)
    q, qd = random_states(model, n, seed=32)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    lim = model.dof_limits()
    for rnd in range(3):
        for o, info in enumerate(model.objects):
            nd = info["n_dofs"]
            if nd == 0:
                continue
            k = int(rng.integers(1, nd + 1))
            idx = np.sort(rng.choice(nd, size=k, replace=False)).astype(np.int32)
            lo = np.maximum(lim[info["dof_offset"] + idx, 0], -2.0)
            hi = np.minimum(lim[info["dof_offset"] + idx, 1], 2.0)
            vals = rng.uniform(lo, hi, (n, k)).astype(np.float32)
            vels = rng.uniform(-3.0, 3.0, (n, k)).astype(np.float32)   # partly outside the velocity limits: clamped
            use_idx = None if k == nd and rnd == 0 else idx
            batch.setDOFPositions(o, vals, use_idx)
            batch.setDOFVelocities(o, vels, use_idx)
            for w, ow in enumerate(worlds):
                ow.set_object_dofs(o, vals[w], use_idx)
                ow.set_object_dofs(o, vels[w], use_idx, velocities=True)
            gp, gv = batch.getDOFPositions(o, use_idx), batch.getDOFVelocities(o, use_idx)
            for w, ow in enumerate(worlds):
                assert np.array_equal(gp[w], ow.get_object_dofs(o, use_idx))
                assert np.array_equal(gv[w], ow.get_object_dofs(o, use_idx, velocities=True))
        assert_same(batch, worlds, "DOF subsets round %d" % rnd)
        batch.stepPhysics(3)
        for ow in worlds:
            ow.step(3)
        assert_same(batch, worlds, "DOF subsets + steps round %d" % rnd)


def test_select_worlds(b2g):
    """b2g_batch_select_worlds: per-world calls on a sub-range touch only that range; stepping refuses a partial selection."""
    n = 40
    batch = make_batch(b2g, ENV_A, n)
    q, qd = random_states(batch.model, n, seed=5)
    batch.setWorldState(q, qd, reset_caches=True)
    before = batch.bodies()
    batch.selectWorlds(33, 4)
    sq, sqd = batch.getWorldState()
    assert sq.shape == (4, batch.nq)
    full_q, _ = (q, qd)
    pose = np.tile(np.float32([0.3, -0.4, 0.5]), (4, 1))
    batch.setObjectPose("obj1", pose)
    assert np.array_equal(batch.getObjectPose("obj1")[:, :2], pose[:, :2])
    with pytest.raises(RuntimeError):
        batch.stepPhysics(1)
    batch.selectWorlds()
    after = batch.bodies()
    same = [np.array_equal(before[w], after[w]) for w in range(n)]
    assert all(same[:33]) and all(same[37:]) and not any(same[33:37])
    # a fresh-cache set_state on a sub-range (broadcast of the loaded world + set_state) leaves the other worlds alone
    batch.selectWorlds(5, 3)
    batch.setWorldState(q[5:8], qd[5:8], reset_caches=True)
    batch.selectWorlds()
    again = batch.bodies()
    assert all(np.array_equal(after[w], again[w]) for w in range(n) if not 5 <= w < 8)
    got_q, _ = batch.getWorldState()
    assert np.array_equal(got_q[33:37], np.concatenate([sq[:, :7], batch.getObjectPose("obj1")[33:37], sq[:, 10:]], 1))
    batch.stepPhysics(2)


def test_control_efforts_parity(b2g):
    """Box2DRobot::setController with a callback returning fixed efforts -> commandEfforts (:2284-2379) on every step;
    switching back to the velocity controller afterwards."""
    n = 48
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=61, spread=0.7)
    rng = np.random.default_rng(62)
    eff = np.concatenate([rng.uniform(-3, 3, (n, 2)), rng.uniform(-0.5, 0.5, (n, 1)), rng.uniform(-0.05, 0.05, (n, 4))], 1).astype(np.float32)
    eff[:, 4][::3] = 0.0     # tau == 0: negative motor speed, zero max torque (:1198-1222)
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setControlEfforts(0, eff)
    batch.stepPhysics(12)
    tg = random_targets(batch.model, 0, n, seed=63)[0]
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_control_efforts(0, eff[w])
        ow.step(12)
    assert_same(batch, worlds, "fixed control efforts, 12 steps")
    batch.setTargetVelocity(0, tg)
    batch.stepPhysics(8)
    for w, ow in enumerate(worlds):
        ow.set_target_velocity(0, tg[w])
        ow.step(8)
    assert_same(batch, worlds, "back to the velocity controller")


def test_reference_test_configurations(b2g):
    """The configurations / velocities of the reference's own test (tests/test_Box2DWorld.cpp:11-20): set then get through
    the DOF API returns them (1e-5) and matches the oracle bit for bit."""
    from test_oracle_kat import REF_OBJECT, REF_ROBOT
    n = 5
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    for obj, conf, vel in (REF_OBJECT, REF_ROBOT):
        batch.setDOFPositions(obj, np.float32(conf))
        batch.setDOFVelocities(obj, np.float32(vel))
        for ow in worlds:
            ow.set_object_dofs(obj, conf)
            ow.set_object_dofs(obj, vel, velocities=True)
        gp, gv = batch.getDOFPositions(obj), batch.getDOFVelocities(obj)
        assert np.allclose(gp, np.float32(conf), atol=1e-5) and np.allclose(gv, np.float32(vel), atol=1e-5)
        for w, ow in enumerate(worlds):
            assert np.array_equal(gp[w], ow.get_object_dofs(obj)) and np.array_equal(gv[w], ow.get_object_dofs(obj, velocities=True))
    assert_same(batch, worlds, "reference test configurations")


def test_one_call_rollout_matches_stepwise(b2g):
    """b2g_batch_rollout (host buffers in / out, one call) == setWorldState + per-segment setTargetVelocity + stepPhysics +
    getWorldState, bit for bit; q = None continues from the current state."""
    n = 300
    batch = make_batch(b2g, ENV_A, n)
    other = make_batch(b2g, ENV_A, n)
    q, qd = random_states(batch.model, n, seed=71, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=72, segments=4)
    oq, oqd = batch.rollout(0, tg[:2], 5, q, qd)
    oq2, oqd2 = batch.rollout(0, tg[2:], 5)                       # continues
    other.setWorldState(q, qd, reset_caches=True)
    for seg in range(4):
        other.setTargetVelocity(0, tg[seg])
        other.stepPhysics(5)
        if seg == 1:
            mq, mqd = other.getWorldState()
            assert np.array_equal(oq, mq) and np.array_equal(oqd, mqd)
    mq, mqd = other.getWorldState()
    assert np.array_equal(oq2, mq) and np.array_equal(oqd2, mqd)
    assert np.array_equal(batch.bodies(), other.bodies())


def test_clone(b2g):
    """Box2DWorld::clone (:3021-3038): same step parameters, DOF state and static object poses, fresh caches: the clone
    continues like an oracle world that went through getWorldState -> fresh world -> setWorldState."""
    n = 40
    batch = make_batch(b2g, ENV_A, n)
    batch.setPhysicsTimeStep(0.005)
    batch.setVelocitySteps(9)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=81, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=82, segments=2)
    batch.setWorldState(q, qd, reset_caches=True)
    scatter_obstacle(batch, worlds, seed=83)          # moves the static obstacle in both
    batch.setTargetVelocity(0, tg[0])
    batch.stepPhysics(15)
    twin = batch.clone()
    assert (twin.getPhysicsTimeStep(), twin.getVelocitySteps(), twin.getPositionSteps()) == (0.005, 9, 15)
    assert np.allclose(twin.getObjectPose("obj2"), batch.getObjectPose("obj2"), atol=2e-6)   # pose -> body -> pose round trip
    cq, cqd = twin.getWorldState()
    bq, bqd = batch.getWorldState()
    # setWorldState clamps joint positions / DOF velocities to their limits like the reference, so only the base poses have
    # to survive the clone unchanged; the full check is the oracle's fresh-world path below
    assert np.allclose(cq[:, [0, 1, 2, 7, 8, 9]], bq[:, [0, 1, 2, 7, 8, 9]], atol=1e-5)
    assert int(twin.contacts()[0].sum()) == 0         # fresh caches
    # oracle: fresh worlds with the same parameters that receive the clone's state
    fresh = oracle_worlds(load_env_dict(ENV_A), n)
    poses = batch.getObjectPose("obj2")
    for w, ow in enumerate(fresh):
        ow.set_params(0.005, 9, 15)
        ow.set_object_pose(2, *[float(x) for x in poses[w]])
        ow.set_state(bq[w], bqd[w])
    assert_same(twin, fresh, "clone == fresh world + setWorldState")
    twin.setTargetVelocity(0, tg[1])
    twin.stepPhysics(10)
    for w, ow in enumerate(fresh):
        ow.set_target_velocity(0, tg[1][w])
        ow.step(10)
    assert_same(twin, fresh, "clone after 10 steps")


def test_cfg1_single_world_thousand_steps(b2g):
    """BASELINE.json configs[0]: test_env.yaml, ONE world, 1000 steps of robot velocity control with a fixed target
    ([0.5, 0, 0.3, 0.5, -0.5, -0.5, 0.5]) from the YAML's initial state -- the longest trajectory in the
    suite; compared with the oracle every 100 steps, bit for bit (plus three neighbours with perturbed targets)."""
    n = 4
    batch = make_batch(b2g, ENV_A, n)
    worlds = oracle_worlds(load_env_dict(ENV_A), n)
    tg = np.tile(np.float32([0.5, 0.0, 0.3, 0.5, -0.5, -0.5, 0.5]), (n, 1))
    tg[1:, :3] += np.float32([[0.3, 0.2, -0.4], [-0.6, 0.1, 0.2], [0.0, -0.5, 0.6]])
    batch.setTargetVelocity(0, tg)
    for w, ow in enumerate(worlds):
        ow.set_target_velocity(0, tg[w])
    for block in range(10):
        batch.stepPhysics(100)
        for ow in worlds:
            ow.step(100)
        assert_same(batch, worlds, "after %d steps" % (100 * (block + 1)))
    q, _ = batch.getWorldState()
    assert abs(q[0, 0] - 5.0) < 0.6 and abs(q[0, 2] - 3.0) < 0.5      # the robot has travelled ~ v * 10 s
    assert np.all(batch.status() == 0)


def _random_env(seed):
    """A random environment within the loader's rules: a robot (base + 0-2 finger chains of 1-2 links, actuated revolute
    joints with limits), 2-4 objects of 1-3 chained links (free hinges, with and without limits), one of them possibly
    static; random convex polygons (boxes, triangles, up to two per link), masses, frictions, restitutions."""
    rng = np.random.default_rng(seed)

    def poly():
        if rng.random() < 0.7:
            hx, hy = rng.uniform(0.05, 0.25), rng.uniform(0.03, 0.12)
            ox, oy = rng.uniform(-0.05, 0.05, 2)
            return [float(v) for v in (ox - hx, oy - hy, ox + hx, oy - hy, ox + hx, oy + hy, ox - hx, oy + hy)]
        r = rng.uniform(0.08, 0.2)
        a = np.sort(rng.uniform(0, 2 * np.pi, 3)) + np.array([0.0, 0.4, 0.8])
        return [float(v) for p in a for v in (r * np.cos(p), r * np.sin(p))]

    def link(name):
        return dict(name=name, geometry=[poly() for _ in range(1 + int(rng.random() < 0.3))], mass=float(rng.uniform(0.1, 3.0)),
                    ground_friction=float(rng.choice([0.0, rng.uniform(0.01, 0.4)])), ground_torque_integral=float(rng.uniform(0.0, 0.2)),
                    contact_friction=float(rng.uniform(0.0, 1.5)), restitution=float(rng.choice([0.0, 0.5, 1.0])), balls=[])

    def joint(name, a, b, actuated):
        lim = [0.0, 0.0] if rng.random() < 0.3 else [float(-rng.uniform(0.3, 1.5)), float(rng.uniform(0.3, 1.5))]
        return dict(name=name, link_a=a, link_b=b, axis=[float(rng.uniform(-0.2, 0.2)), float(rng.uniform(0.05, 0.3))],
                    axis_orientation=float(rng.uniform(-0.5, 0.5)), position_limits=lim, velocity_limits=[-10.0, 10.0],
                    acceleration_limits=[-30.0, 30.0], joint_type="revolute", actuated=actuated)

    def chain_object(name, n_chains, max_len, actuated):
        links, joints = [link("base")], []
        for ch in range(n_chains):
            parent = "base"
            for k in range(int(rng.integers(1, max_len + 1))):
                nm = "l%d_%d" % (ch, k)
                links.append(link(nm))
                joints.append(joint("j%d_%d" % (ch, k), parent, nm, actuated))
                parent = nm
        return dict(name=name, static=False, links=links, joints=joints, base_actuation=None)

    robot = chain_object("robot", int(rng.integers(0, 3)), 2, True)
    robot["base_actuation"] = dict(translational_velocity_limits=[-2.0, 2.0], translational_acceleration_limits=[-3.0, 3.0],
                                   rotational_velocity_limits=[-2.0, 2.0], rotational_acceleration_limits=[-2.0, 2.0], use_center_of_mass=True)
    objects = [chain_object("obj%d" % i, int(rng.random() < 0.4), 2, False) for i in range(int(rng.integers(2, 5)))]
    if rng.random() < 0.6:
        objects[-1]["static"] = True
    return dict(scale=10.0, world_bounds=[-10, -10, 10, 10], robots=[robot], objects=objects, states={})


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_random_models_parity(b2g, seed):
    """The engine is generic over the model: random robots / articulated objects / static obstacles built through the
    programmatic environment API, scattered so that they collide, stepped with random targets -- bit-exact vs the oracle."""
    env = _random_env(seed)
    n = 24
    model = b2g.Model(env)
    batch = b2g.BatchBox2DWorld(n, max_contacts=96).loadWorld(model)
    worlds = oracle_worlds(env, n)
    assert_same(batch, worlds, "fresh random world %d" % seed)
    q, qd = random_states(model, n, seed=seed + 1000, spread=0.8, vel=1.5)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    rng = np.random.default_rng(seed + 2000)
    for o, info in enumerate(model.objects):          # static objects are placed by pose
        if info["is_static"]:
            poses = np.concatenate([rng.uniform(-0.8, 0.8, (n, 2)), rng.uniform(-3, 3, (n, 1))], 1).astype(np.float32)
            batch.setObjectPose(o, poses)
            for w, ow in enumerate(worlds):
                ow.set_object_pose(o, *[float(x) for x in poses[w]])
    assert_same(batch, worlds, "random world %d after set_state" % seed)
    tg = random_targets(model, 0, n, seed=seed + 3000, segments=4)
    touching = 0
    for seg in range(4):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "random world %d, segment %d" % (seed, seg))
        touching += int(batch.contacts(max_per_world=96)[1][..., 2].sum())
    assert touching > 10, "random scenario %d produced too few touching contacts" % seed
    assert np.all(batch.status() == 0)


@pytest.mark.parametrize("seed", [3, 10, 21, 32])
def test_random_api_sequences(b2g, seed):
    """Random sequences of 40 API calls (stepping with and without sleeping, targets beyond the limits, control efforts, DOF
    subsets, poses, per-world link enable / disable, link mass / ground friction edits, checkCollision, region queries)
    mirrored on oracle worlds; everything is compared after every call.  tools/fuzz_api.py runs the same over hundreds of
    seeds (1700 sequences so far, 0 failures)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_api", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                          "tools", "fuzz_api.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except SystemExit:
        pass
    mod.run(seed, [ENV_A, ENV_B, ENV_2, ENV_3][seed % 4])


@pytest.mark.parametrize("env_name", [ENV_A, ENV_2])
def test_link_properties_parity(b2g, env_name):
    """Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction
    (reference :478-547) on every world of a batch: the model records the kernels read (mass data, joint inverse masses,
    motor / angular masses, ground-friction limits, fixture friction) and b2Body::SetMassData's sweep-centre update leave
    the worlds bit-identical to oracle worlds that received the same setter calls, right away and after stepping on."""
    n = 48
    batch, worlds = scattered_start(b2g, env_name, n, seed=141)
    robot_base = batch.model.objects[0]["first_body"]
    obj1 = batch.model.objects[1]["first_body"]
    obj2 = batch.model.objects[2]["first_body"]

    def both(body, prop, value):
        batch.setLinkProperty(body, prop, value)
        for ow in worlds:
            ow.set_link_property(body, prop, value)
        assert batch.getLinkProperty(body, prop) == worlds[0].get_link_property(body, prop)

    # contact friction before any contact exists
    both(obj1, "contact_friction", 1.3)
    both(robot_base + 3, "contact_friction", 0.02)
    batch.stepPhysics(12)
    for ow in worlds:
        ow.step(12)
    assert_same(batch, worlds, "contact friction set before stepping")
    both(obj1, "mass", 0.21)
    both(robot_base, "mass", 1.7)
    both(robot_base + 2, "mass", 0.05)
    both(obj2, "mass", 2.0)                       # static in test_env.yaml (no-op on the body), dynamic in env2
    both(obj1, "ground_friction", 0.45)
    both(obj1, "ground_torque_integral", 0.03)
    both(robot_base + 1, "ground_friction", 0.0)
    assert_same(batch, worlds, "right after the setters (sweep centres, velocities)")
    tg = random_targets(batch.model, 0, n, seed=143)[0]
    batch.setTargetVelocity(0, tg)
    for w, ow in enumerate(worlds):
        ow.set_target_velocity(0, tg[w])
    batch.stepPhysics(30)
    for ow in worlds:
        ow.step(30)
    assert_same(batch, worlds, "30 steps after the setters")
    assert int(batch.contacts()[0].sum()) > 10, "scenario has too few contacts"
    # contact friction with contacts alive: like upstream (b2Contact mixes the coefficients when it is created) they keep
    # the old value until they end, contacts that begin later use the new one
    touching_before = int(batch.contacts()[1][..., 2].sum())
    both(obj1, "contact_friction", 0.01)
    both(batch.model.n_bodies - 1, "contact_friction", 1.9)   # the last link (env2's robot is one rigid link)
    both(robot_base, "contact_friction", 0.6)
    batch.stepPhysics(25)
    for ow in worlds:
        ow.step(25)
    assert_same(batch, worlds, "25 steps after changing contact friction under live contacts")
    assert touching_before > 0
    # the clone reloads the environment description (reference :3030): it does not inherit the changed parameters
    twin = batch.clone()
    assert twin.getLinkProperty(obj1, "mass") == batch.model.get_link_property(obj1, "mass") != batch.getLinkProperty(obj1, "mass")
    # per-world selection cannot address a parameter of the shared model
    batch.selectWorlds(3, 5)
    with pytest.raises(b2g.B2GError):
        batch.setLinkProperty(obj1, "mass", 1.0)
    batch.selectWorlds(0, 0)


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2])
def test_link_states_parity(b2g, env_name):
    """Box2DLink::getPose / getVelocityVector / getCenterOfMass (reference :215-241, 595-603) of every link of every world,
    bit-identical to the oracle's links after a contact-rich rollout; also on a selected sub-range of worlds."""
    n = 40
    batch, worlds = scattered_start(b2g, env_name, n, seed=151)
    batch.stepPhysics(30)
    for ow in worlds:
        ow.step(30)
    ls = batch.getLinkStates()
    assert ls.shape == (n, batch.model.n_bodies, 8)
    for w, ow in enumerate(worlds):
        assert np.array_equal(ls[w], ow.link_states()), "world %d" % w
    assert np.abs(ls[:, 1:, 3:6]).max() > 0.01                 # something moves
    batch.selectWorlds(7, 5)
    assert np.array_equal(batch.getLinkStates(), ls[7:12])
    batch.selectWorlds(0, 0)
    # base link pose == the object's pose DOFs (what getWorldState reports)
    q, _ = batch.getWorldState()
    for o in batch.model.objects:
        if not o["is_static"]:
            base = [b for b in range(o["first_body"], o["first_body"] + o["n_links"]) if batch.model.link(b)["is_base_link"]][0]
            assert np.array_equal(ls[:, base, :3], q[:, o["dof_offset"]:o["dof_offset"] + 3])


# ---- round 2 ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_name", [ENV_A, ENV_B, ENV_2, ENV_3, "clutter"])
def test_single_step_vs_libm_trig_oracle(b2g, env_name):
    """North-star tolerance against reference-like trigonometry: the oracle rebuilt with libm sinf / cosf in b2Rot (what
    upstream Box2D calls; the CUDA path and the oracle proper share one fixed sin / cos sequence, so their difference is 0 by
    construction).  One step from an identical, physically feasible fresh-cache state: |dq|, |dqd| <= 1e-5 (m, rad, m/s,
    rad/s); the touching contact-pair sets with their manifold point counts identical.  (From interpenetrating random states
    the same comparison stays within 7e-6 on positions but reaches 7e-5 rad/s on a few percent of the worlds: the dynamics
    amplify a last-bit difference of sin / cos there)"""
    from oracle import b2o
    lm = b2o.load_variant("_libm")
    n = 192
    if env_name == "clutter":
        from box2d_sim_env_b200 import envgen
        desc = envgen.clutter_env()
        batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.Model(desc))
        q, qd = envgen.clutter_states(n, seed=31, speed=2.0)
        # a second step target: let the clutter slide into contact first (60 steps on both sides with the exact oracle), then
        # compare ONE step of the libm variant from that state
        warm = 60
    else:
        desc = env_dict(env_name)
        batch = make_batch(b2g, env_name, n)
        q, qd = feasible_states(batch.model, n, 31, base_front_edge(batch.fat_aabbs()[0]))
        warm = 0
    tgt = random_targets(batch.model, 0, n, seed=32)[0]
    tgt[:, 1] = np.abs(tgt[:, 1])   # the robot pushes towards the box
    if warm:
        ob = b2o.OracleBatch(b2o.OracleEnv(desc), n)
        q, qd = ob.rollout(q, qd, 0, tgt[None], warm, os.cpu_count() or 4)
    gq, gqd = batch.rollout(0, tgt[None], 1, q, qd, reset_caches=True)
    counts, cio, _ = batch.contacts(max_per_world=128)
    ol = lm.OracleBatch(lm.OracleEnv(desc), n)
    lq, lqd = ol.rollout(q, qd, 0, tgt[None], 1, os.cpu_count() or 4)
    assert np.max(np.abs(gq.astype(np.float64) - lq)) <= 1e-5, np.max(np.abs(gq - lq))
    assert np.max(np.abs(gqd.astype(np.float64) - lqd)) <= 1e-5, np.max(np.abs(gqd - lqd))
    touching = 0
    for w in range(n):
        oio, _ = ol.world(w).contacts()
        g = sorted((int(r[0]), int(r[1]), int(r[4])) for r in cio[w, :int(counts[w])] if r[2])
        o = sorted((int(r[0]), int(r[1]), int(r[4])) for r in oio if r[2])
        assert g == o, "world %d: touching pair sets differ between the CUDA path and the libm-trig oracle" % w
        touching += len(o)
    assert touching > 10, "scenario produced too few touching contacts"


@pytest.mark.parametrize("lpw,lanes", [(4, None), (2, None), (8, 3), (16, None), (32, None)])
@pytest.mark.parametrize("env_name", [ENV_A, ENV_B])
def test_sub_warp_per_world_kernel_parity(b2g, env_name, lpw, lanes, monkeypatch):
    """k_step_multi: a world owned by 2 .. 32 lanes (lane-strided body / joint / contact loops, dataflow-scheduled velocity
    iterations from the cached slot table).  Contact-rich 120-step rollout, every segment bit-identical to the oracle; the
    schedule cache is exercised by topologies that change as contacts begin and end."""
    monkeypatch.setenv("B2G_LPW", str(lpw))
    if lanes:
        monkeypatch.setenv("B2G_LANES", str(lanes))
    n = 77
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    q, qd = random_states(batch.model, n, seed=41, spread=0.5)
    tg = random_targets(batch.model, 0, n, seed=42, segments=12)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    scatter_obstacle(batch, worlds, seed=43)
    touching = 0
    for seg in range(12):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10, allow_sleeping=(seg % 4 != 3))
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10, True, seg % 4 != 3)
        assert_same(batch, worlds, "lpw %d segment %d" % (lpw, seg))
        touching += int(batch.contacts()[1][..., 2].sum())
    assert touching > 50
    assert np.all(batch.status() == 0)


def test_sub_warp_per_world_clutter_and_serial_fallback(b2g, monkeypatch):
    """Shape C on the sub-warp kernel: dozens of simple islands strided over the lanes, merged islands whose constraint count
    exceeds the schedule's 32-node limit take the sequential fallback; 2 velocity iterations exercise short lap counts."""
    from box2d_sim_env_b200 import envgen
    monkeypatch.setenv("B2G_LPW", "4")
    n = 20
    env = envgen.clutter_env()
    batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.Model(env))
    worlds = oracle_worlds(env, n)
    q, qd = envgen.clutter_states(n, seed=8, speed=3.0)
    tg = random_targets(batch.model, 0, n, seed=9, segments=5)
    batch.setWorldState(q, qd, reset_caches=True)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
    for seg in range(5):
        if seg == 3:
            batch.setVelocitySteps(2)
            batch.setPositionSteps(3)
            for ow in worlds:
                ow.set_params(0.01, 2, 3)
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "clutter on 4 lanes per world, segment %d" % seg)
    assert np.all(batch.status() == 0)


def test_pinned_empty_and_current_device(b2g):
    """pinned_empty returns page-locked memory that the one-call rollout reads and writes directly and that is released with
    the array; entry points leave the caller's current CUDA device alone."""
    import gc
    import torch
    n = 64
    batch = make_batch(b2g, ENV_A, n)
    q = b2g.pinned_empty((n, batch.nq))
    qd = b2g.pinned_empty((n, batch.nq))
    q0, qd0 = random_states(batch.model, n, seed=51, spread=0.6)
    q[:] = q0
    qd[:] = qd0
    tg = b2g.pinned_empty((2, n, 7))
    tg[:] = random_targets(batch.model, 0, n, seed=52, segments=2)
    oq, oqd = b2g.pinned_empty((n, batch.nq)), b2g.pinned_empty((n, batch.nq))
    batch.rollout_host(0, q.ctypes.data, qd.ctypes.data, tg.ctypes.data, 2, 5, oq.ctypes.data, oqd.ctypes.data)
    gq, gqd = batch.getWorldState()
    assert np.array_equal(gq, oq) and np.array_equal(gqd, oqd)
    view = oq[:4]
    del oq
    gc.collect()
    assert np.array_equal(view, gq[:4])   # the block lives as long as any view of it
    if torch.cuda.device_count() > 1:
        other = b2g.BatchBox2DWorld(8, device=1).loadWorld(b2g.data_path(ENV_A))
        torch.cuda.set_device(0)
        other.stepPhysics(1)
        assert torch.cuda.current_device() == 0


def test_multi_gpu_c_abi_rollout(b2g):
    """b2g_multi_*: the batch sharded over the GPUs of the node inside one process, ncclAllGather of the final states.  The
    result must equal the single-batch rollout bit for bit, whatever the device count and however ragged the slices (runs
    with one device on a single-GPU box, with two when the box has them)."""
    import torch
    ndev = min(2, torch.cuda.device_count())
    n = 1001   # odd: slices of 501 + 500 worlds on two devices
    model = b2g.Model(b2g.data_path(ENV_A))
    multi = b2g.MultiGpuBatch(model, n, list(range(ndev)))
    single = b2g.BatchBox2DWorld(n).loadWorld(model)
    q, qd = random_states(model, n, seed=61, spread=0.6)
    tg = random_targets(model, 0, n, seed=62, segments=3)
    mq, mqd = multi.rollout(0, tg, 10, q, qd)
    sq, sqd = single.rollout(0, tg, 10, q, qd)
    assert np.array_equal(mq, sq) and np.array_equal(mqd, sqd)
    sizes = [multi.shard(i)[2] for i in range(ndev)]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1 and multi.shard(0)[1] == 0
    # the gathered copy every device holds: rows of device 0's slice on the last device
    qa, qda, rows = multi.gathered_dev(ndev - 1)
    t = torch.empty((ndev, rows, model.nq), dtype=torch.float32, device="cuda:%d" % (ndev - 1))
    import ctypes
    from box2d_sim_env_b200.batch_world import lib as _lib  # noqa: F401
    torch.cuda.synchronize()
    ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(qa), ctypes.c_size_t(t.numel() * 4), 3)
    assert np.array_equal(t[0, :sizes[0]].cpu().numpy(), sq[:sizes[0]])
    tm = multi.last_timing()
    assert tm["kernel_ms_max"] >= tm["kernel_ms_min"] > 0.0 and tm["gather_ms"] > 0.0
    # a second call reuses the communicators; continuing without set_state matches as well
    mq2, _ = multi.rollout(0, tg[:1], 5)
    sq2, _ = single.rollout(0, tg[:1], 5)
    assert np.array_equal(mq2, sq2)
    multi.close()


@pytest.mark.parametrize("variant", [dict(), dict(actuated=True), dict(limits=(0.1, 0.2)), dict(limits=(0.05, 0.0501))])
@pytest.mark.parametrize("lpw", [1, 4])
def test_prismatic_joint_parity(b2g, variant, lpw, monkeypatch):
    """Row 8f-4: b2PrismaticJoint (init / velocity / position constraints, lower / upper / equal limits, the motor of an
    actuated joint as a brake) on both stepping kernels.  The robot, placed per world through its own DOF setters, rams a
    drawer whose slider runs on a prismatic rail; bodies, joints (impulses, limit states) and contacts bit-identical to the
    oracle after every 10-step segment; getWorldState reports the rail's translation and speed in Box2D units like the
    reference.  Setters on the drawer are refused (the reference throws there)."""
    from parity_utils import env_prismatic
    if lpw > 1:
        monkeypatch.setenv("B2G_LPW", str(lpw))
    n = 48
    env = env_prismatic(**variant)
    batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.Model(env))
    worlds = oracle_worlds(env, n)
    assert_same(batch, worlds, "fresh prismatic world")
    rng = np.random.default_rng(71)
    rq = np.zeros((n, 7), np.float32)
    rq[:, 0] = -0.5 + rng.uniform(-0.05, 0.05, n)
    rq[:, 1] = rng.uniform(-0.08, 0.08, n)
    rq[:, 2] = -1.5707964 + rng.uniform(-0.2, 0.2, n)
    rq[:, 3:] = rng.uniform(-0.5, 0.5, (n, 4))
    batch.setDOFPositions("my_robot", rq)
    for w, ow in enumerate(worlds):
        ow.set_object_dofs(0, rq[w])
    tg = random_targets(batch.model, 0, n, seed=72, segments=12)
    tg[:, :, 0] = np.abs(tg[:, :, 0])      # towards the drawer
    states = set()
    for seg in range(12):
        batch.setTargetVelocity(0, tg[seg])
        batch.stepPhysics(10)
        for w, ow in enumerate(worlds):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(10)
        assert_same(batch, worlds, "prismatic segment %d" % seg)
        gq, gqd = batch.getWorldState()
        oq = np.stack([ow.get_state()[0] for ow in worlds])
        oqd = np.stack([ow.get_state()[1] for ow in worlds])
        assert np.array_equal(gq, oq) and np.array_equal(gqd, oqd)
        states |= set(int(x) for x in batch.joints()[0][:, -1, 3])
    assert int(batch.contacts()[1][..., 2].sum()) > 0 or len(states) > 1
    if "limits" in variant:
        assert states & {1, 2, 3}, "no limit state was exercised: %s" % states
    q, qd = batch.getWorldState()
    for call in (lambda: batch.setWorldState(q, qd), lambda: batch.setDOFPositions("drawer", q[:, 7:]), lambda: batch.setToRest()):
        with pytest.raises(b2g.B2GError) as e:
            call()
        assert e.value.code == 4


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B])
def test_position_controller_parity_and_convergence(b2g, env_name):
    """Row 8f-4, second half: the position controller on top of the velocity controller (SE2 controller for the rigid robot of
    shape B, all-DOF controller for shape A; Box2DWorldViewer.cpp:1208-1216 instantiates sim_env's classes, whose law is assumed:
    v_target = kp * (q_target - q) with the heading error wrapped).  Bit-identical to the oracle's implementation of the same
    law; first-principles check: worlds whose way is free reach their targets, the heading through the short way round."""
    n = 96
    batch = make_batch(b2g, env_name, n)
    worlds = oracle_worlds(env_dict(env_name), n)
    nd = batch.model.objects[0]["n_dofs"]
    rng = np.random.default_rng(81)
    q, qd = batch.getWorldState()
    q[:, nd] = 5.0 + rng.uniform(-0.5, 0.5, n)       # obj1 out of the way in most worlds ...
    q[: n // 4, nd:nd + 2] = [0.0, 0.45]              # ... in the robot's path in a quarter of them
    q[:, 2] = rng.uniform(-3.0, 3.0, n)
    batch.setWorldState(q, qd, reset_caches=True)
    tgt = np.zeros((n, nd), np.float32)
    tgt[:, 0:2] = rng.uniform(-0.6, 0.6, (n, 2))
    tgt[:, 2] = rng.uniform(-3.0, 3.0, n)
    tgt[:, 3:] = rng.uniform(-0.8, 0.8, (n, nd - 3))
    batch.setTargetPosition(0, tgt, kp=4.0)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_target_position(0, tgt[w], 4.0)
    for seg in range(10):
        batch.stepPhysics(30)
        for ow in worlds:
            ow.step(30)
        assert_same(batch, worlds, "position control, segment %d" % seg)
    gq, _ = batch.getWorldState()
    err = gq[n // 4:, :nd] - tgt[n // 4:]
    err[:, 2] = (err[:, 2] + np.pi) % (2 * np.pi) - np.pi
    err0 = q[n // 4:, :nd] - tgt[n // 4:]
    err0[:, 2] = (err0[:, 2] + np.pi) % (2 * np.pi) - np.pi
    # the base reaches its x, y target; heading and fingers stop where the velocity controller's effort (inertia x the
    # acceleration limit) no longer beats the ground friction -- closer than they started, through the short way round
    assert np.abs(err[:, :2]).max() < 0.05, np.abs(err).max(axis=0)
    assert np.all(np.abs(err[:, 2]) <= np.abs(err0[:, 2]) + 0.05) and np.abs(err[:, 2]).max() < 0.6, np.abs(err).max(axis=0)
    # (finger targets that make the two fingers collide cannot be reached: no bound on the joint errors)
    assert np.median(np.abs(err[:, 3:])) < 0.2 if nd > 3 else True
    # switching back to velocity control leaves position mode
    batch.setTargetVelocity(0, np.zeros((n, nd), np.float32))
    for ow in worlds:
        ow.set_target_velocity(0, np.zeros(nd, np.float32))
    batch.stepPhysics(20)
    for ow in worlds:
        ow.step(20)
    assert_same(batch, worlds, "back to velocity control")


def test_reset_caches_keeps_static_object_poses(b2g):
    """reset_caches = what Box2DWorld::clone() gives (:3026-3037: a freshly loaded world that receives the world state): the
    poses given to static objects survive -- through the same getPose -> setPose hand-over a clone does --, contacts / impulses /
    sleep timers / targets do not."""
    from oracle import b2o
    n = 64
    batch = make_batch(b2g, ENV_A, n)
    old = oracle_worlds(env_dict(ENV_A), n)
    q, qd = random_states(batch.model, n, seed=91, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=92)[0]
    # dirty the worlds: contacts, impulses, a target, a moved obstacle
    batch.setWorldState(q, qd, reset_caches=True)
    scatter_obstacle(batch, old, seed=93)
    batch.setTargetVelocity(0, tg)
    batch.stepPhysics(25)
    # a planner's next query: fresh caches, new state; the obstacle stays where it was put
    q2, qd2 = random_states(batch.model, n, seed=94, spread=0.6)
    batch.setWorldState(q2, qd2, reset_caches=True)
    env = b2o.OracleEnv(env_dict(ENV_A))
    worlds = []
    for w, ow in enumerate(old):                      # the oracle's clone(): new world, static poses handed over, then the state
        fresh = b2o.OracleWorld(env)
        fresh.set_object_pose(2, *[float(x) for x in ow.get_object_pose(2)])
        fresh.set_state(q2[w], qd2[w])
        worlds.append(fresh)
    assert_same(batch, worlds, "after reset_caches")
    batch.stepPhysics(30)
    for ow in worlds:
        ow.step(30)
    assert_same(batch, worlds, "30 steps after reset_caches")
    assert int(batch.contacts()[1][..., 2].sum()) > 10


@pytest.mark.parametrize("mode", ["narrow", "full-warps"])
def test_link_parameter_sets_per_tile(b2g, mode, monkeypatch):
    """Per-world link parameters (reference Box2DLink::setMass / setGroundFrictionCoefficient /
    setContactFriction :478-547 called on the links of DIFFERENT worlds): a tile-aligned selection of worlds gets its own
    parameter set, every warp stages the blob of its tile.  Worlds of four tiles with four different parameter histories
    stay bit-identical to oracle worlds that received the same setter calls, on the narrow-warp and on the CTA-synchronised
    stepping kernel, through stepping, checkCollision (contacts created there mix the tile's friction) and one-world views."""
    if mode == "full-warps":
        monkeypatch.setenv("B2G_LANES", "32")
        monkeypatch.setenv("B2G_STEP_BLOCK", "64")
    n = 100   # tiles: 32 + 32 + 32 + 4
    batch, worlds = scattered_start(b2g, ENV_A, n, seed=311)
    robot_base = batch.model.objects[0]["first_body"]
    obj1 = batch.model.objects[1]["first_body"]

    def on(first, count, body, prop, value):
        batch.selectWorlds(first, count)
        batch.setLinkProperty(body, prop, value)
        batch.selectWorlds(0, 0)
        for ow in worlds[first:first + count]:
            ow.set_link_property(body, prop, value)

    on(32, 32, obj1, "mass", 0.31)
    on(64, 36, obj1, "ground_friction", 0.55)
    on(0, 32, robot_base, "mass", 2.2)
    on(32, 68, obj1, "contact_friction", 1.4)           # splits nothing new for tile 1, patches the sets of tiles 1-3
    on(64, 32, robot_base + 2, "mass", 0.07)             # tile 2 leaves the set it shared with tile 3
    assert_same(batch, worlds, "right after the per-tile setters")
    batch.stepPhysics(25)
    for ow in worlds:
        ow.step(25)
    assert_same(batch, worlds, "25 steps with four parameter sets")
    # a setter on the whole batch reaches every set
    batch.setLinkProperty(obj1, "ground_torque_integral", 0.02)
    for ow in worlds:
        ow.set_link_property(obj1, "ground_torque_integral", 0.02)
    tg = random_targets(batch.model, 0, n, seed=313)[0]
    batch.setTargetVelocity(0, tg)
    for w, ow in enumerate(worlds):
        ow.set_target_velocity(0, tg[w])
    batch.stepPhysics(30)
    for ow in worlds:
        ow.step(30)
    assert_same(batch, worlds, "30 more steps after a whole-batch setter")
    assert int(batch.contacts()[1][..., 2].sum()) > 10, "scenario has too few touching contacts"
    # the getters answer for the first selected world's tile
    batch.selectWorlds(40, 1)
    assert batch.getLinkProperty(obj1, "mass") == worlds[40].get_link_property(obj1, "mass")
    batch.selectWorlds(70, 1)
    assert batch.getLinkProperty(obj1, "ground_friction") == worlds[70].get_link_property(obj1, "ground_friction")
    # per-world calls on a selection that does not start at a tile boundary still read the right parameter set
    batch.selectWorlds(45, 30)   # worlds 45..74: tiles 1 and 2
    q, qd = batch.getWorldState()
    for i, w in enumerate(range(45, 75)):
        oq, oqd = worlds[w].get_state()
        assert np.array_equal(q[i], oq) and np.array_equal(qd[i], oqd)
    # a selection inside a tile cannot carry its own parameters
    batch.selectWorlds(33, 10)
    with pytest.raises(b2g.B2GError):
        batch.setLinkProperty(obj1, "mass", 1.0)
    batch.selectWorlds(0, 0)
    counts, cio, cfo = batch.checkCollision(max_per_world=32)
    for w, ow in enumerate(worlds):
        oio, ofo = ow.check_collision()
        assert counts[w] == len(oio) and np.array_equal(cio[w, :len(oio)], oio), "world %d: checkCollision differs" % w
    batch.stepPhysics(10)
    for ow in worlds:
        ow.step(10)
    assert_same(batch, worlds, "after checkCollision and 10 more steps")


def test_status_marks_scaled_targets(b2g):
    """the rule that scales a target beyond the velocity limits (sim_env's scaleToLimits) is an assumption;
    worlds whose results depend on it carry bit 2 in their status word (a note: b2g_batch_status still returns B2G_OK)."""
    n = 8
    batch = make_batch(b2g, ENV_A, n)
    lim = batch.model.dof_limits()[:7, 2:4]
    tg = np.zeros((n, 7), np.float32)
    tg[:] = 0.5 * lim[:, 1]
    tg[3, 0] = 3.0 * lim[0, 1]          # world 3: base x beyond the limit
    tg[6, 5] = 2.0 * lim[5, 0]          # world 6: a finger joint below its lower limit
    batch.setTargetVelocity(0, tg)
    batch.stepPhysics(2)
    assert batch.status().tolist() == [0, 0, 0, 4, 0, 0, 4, 0]
    q, qd = batch.getWorldState()
    batch.setWorldState(q, qd, reset_caches=True)   # a fresh clone forgets it
    assert np.all(batch.status() == 0)


@pytest.mark.parametrize("env_name,n,group", [(ENV_A, 40000, None), (ENV_B, 33333, "0"), (ENV_A, 96001, "4096")])
def test_regrouped_rollout_parity(b2g, env_name, n, group, monkeypatch):
    """Opt-in regrouping (B2G_REGROUP=1): large rollouts sort the worlds by island shape into a second state array before
    every target segment -- inside the slot range of one CTA, of the whole batch or of a given group -- and move them back at
    the end (b2g_capi.cpp: regroupedRollout).  Final states of all worlds, and bodies / joints / contact lists of a sample,
    are bit-identical to the oracle; stepping on afterwards works on slot == world again."""
    from oracle import b2o
    monkeypatch.setenv("B2G_REGROUP", "1")
    if group is not None:
        monkeypatch.setenv("B2G_REGROUP_GROUP", group)
    batch = make_batch(b2g, env_name, n)
    q, qd = random_states(batch.model, n, seed=201, spread=0.6)
    tg = random_targets(batch.model, 0, n, seed=202, segments=5)
    gq, gqd = batch.rollout(0, tg, 8, q, qd, reset_caches=True)
    ob = b2o.OracleBatch(env_dict(env_name), n)
    oq, oqd = ob.rollout(q, qd, 0, tg, 8, os.cpu_count() or 4)
    bad = np.nonzero(~(np.all(gq == oq, axis=1) & np.all(gqd == oqd, axis=1)))[0]
    assert bad.size == 0, "%d of %d worlds differ, first %s" % (bad.size, n, bad[:5].tolist())
    m = 300
    worlds = [ob.world(w) for w in range(m)]
    batch.selectWorlds(0, m)
    assert_same(batch, worlds, "sample after the regrouped rollout")
    batch.selectWorlds(0, 0)
    assert int(batch.contacts(max_per_world=8)[1][..., 2].sum()) > 100, "scenario has too few touching contacts"
    # stepping on without targets (plain launch, slot == world)
    batch.stepPhysics(7)
    for ow in worlds:
        ow.step(7)
    batch.selectWorlds(0, m)
    assert_same(batch, worlds, "7 plain steps after the regrouped rollout")
    batch.selectWorlds(0, 0)
    assert np.all((batch.status() & 3) == 0)
