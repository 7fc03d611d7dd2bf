"""Ball approximation: Box2DLinkDescription::balls (Box2DIOUtils.h:21, 139-142), Box2DLink::_balls
(Box2DWorld.cpp:54-57), getLocalBallApproximation (:727-731), Box2DObject::getBallApproximation (:1754-1763).
CPU: both loaders read the key, the oracle against first-principles known answers.  GPU: CUDA vs oracle, bit-exact."""
import os
import shutil

import numpy as np
import pytest

from parity_utils import ENV_A, load_env_dict, oracle_worlds, random_states, random_targets


def env_with_balls():
    env = load_env_dict(ENV_A)
    rng = np.random.default_rng(5)
    for obj in env["robots"] + env["objects"]:
        for k, link in enumerate(obj["links"]):
            n = 1 + (k % 3)
            link["balls"] = [[float(x) for x in rng.uniform(-0.2, 0.2, 3)] + [float(rng.uniform(0.02, 0.1))] for _ in range(n)]
    return env


def expected_order(env_obj):
    """Box2DObject::getBallApproximation walks its std::map of links: name order."""
    return [b for link in sorted(env_obj["links"], key=lambda l: l["name"]) for b in link["balls"]]


def test_yaml_loaders_read_balls(b2g, oracle, tmp_path):
    data = os.path.dirname(b2g.data_path(ENV_A))
    shutil.copytree(data, tmp_path / "data")
    obj = tmp_path / "data" / "objects" / "test_object.yaml"
    text = obj.read_text().replace("    mass:", "    balls: [[[0.1, -0.2, 0.0], 0.05], [[0.0, 0.3, 0.25], 0.125]]\n    mass:", 1)
    assert "balls" in text
    obj.write_text(text)
    path = str(tmp_path / "data" / ENV_A)
    env = oracle.load_env_yaml(path)                      # PyYAML reader (oracle side)
    m = b2g.Model(path)                                   # product C++ reader
    i = m.object_index("obj1")
    assert m.objects[i]["n_balls"] == 2 and sum(o["n_balls"] for o in m.objects) == 2
    bodies, xyzr = m.local_ball_approximation(i)
    assert bodies.tolist() == [m.objects[i]["first_body"]] * 2
    want = np.array([[0.1, -0.2, 0.0, 0.05], [0.0, 0.3, 0.25, 0.125]], np.float32)
    assert np.array_equal(xyzr, want)
    assert np.array_equal(np.array(env["objects"][0]["links"][0]["balls"], np.float32), want)
    w = oracle.OracleWorld(env)
    assert w.ball_approximation(i).shape == (2, 4)


def test_model_ball_order_and_counts(b2g):
    env = env_with_balls()
    m = b2g.Model(env)
    descs = env["robots"] + env["objects"]
    for i, d in enumerate(descs):
        want = np.array(expected_order(d), np.float32).reshape(-1, 4)
        bodies, xyzr = m.local_ball_approximation(i)
        assert m.objects[i]["n_balls"] == len(want)
        assert np.array_equal(xyzr, want)
        names = sorted(l["name"] for l in d["links"])
        order = [l["name"] for l in d["links"]]
        body_of = {n: m.objects[i]["first_body"] + order.index(n) for n in names}
        assert bodies.tolist() == [body_of[l["name"]] for l in sorted(d["links"], key=lambda l: l["name"]) for _ in l["balls"]]


def test_oracle_ball_known_answers(oracle):
    """First principles in double precision: centre = R(theta) * c + link origin; the radius passes through."""
    env = env_with_balls()
    w = oracle.OracleWorld(env)
    q, qd = w.get_state()
    q[7:10] = [0.4, -0.3, 0.7]     # obj1 pose (base link origin = centre of mass)
    w.set_state(q, np.zeros_like(qd))
    got = w.ball_approximation(1)
    want = np.array(expected_order(env["objects"][0]), np.float64)
    c, s = np.cos(0.7), np.sin(0.7)
    ex = np.stack([c * want[:, 0] - s * want[:, 1] + 0.4, s * want[:, 0] + c * want[:, 1] - 0.3, want[:, 2], want[:, 3]], 1)
    assert np.allclose(got, ex, atol=2e-6)
    assert np.array_equal(got[:, 3], want[:, 3].astype(np.float32))


@pytest.mark.gpu
def test_ball_approximation_parity(b2g):
    env = env_with_balls()
    n = 48
    model = b2g.Model(env)
    batch = b2g.BatchBox2DWorld(n).loadWorld(model)
    worlds = oracle_worlds(env, n)
    q, qd = random_states(model, n, seed=21)
    tg = random_targets(model, 0, n, seed=22)[0]
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, tg)
    batch.stepPhysics(7)
    for w, ow in enumerate(worlds):
        ow.set_state(q[w], qd[w])
        ow.set_target_velocity(0, tg[w])
        ow.step(7)
    for o in range(model.n_objects):
        got = batch.getBallApproximation(o)
        assert got.shape == (n, model.objects[o]["n_balls"], 4)
        for w, ow in enumerate(worlds):
            assert np.array_equal(got[w], ow.ball_approximation(o)), "object %d world %d" % (o, w)
