"""Frozen trajectories (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).  CPU: the oracle still
reproduces them bit for bit.  GPU: the CUDA path reproduces them bit for bit through the C ABI — without the oracle."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

NAMES = sorted(make_golden.CASES)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_fixture_inputs_are_reproducible(name):
    g = load(name)
    _, _, q, qd, tg, steps = make_golden.case_inputs(name)
    assert np.array_equal(g["q0"], q) and np.array_equal(g["qd0"], qd) and np.array_equal(g["targets"], tg)
    assert int(g["steps"]) == steps


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle, name):
    g = load(name)
    now = make_golden.run_oracle(name)
    for key in ("q", "qd", "bodies", "n_contacts", "n_touching", "traj_q"):
        assert np.array_equal(g[key], now[key]), "%s: %s drifted from the frozen fixture" % (name, key)


def test_golden_cases_exercise_contacts():
    assert int(load("shape_a")["n_touching"].sum()) > 0
    assert int(load("shape_c")["n_contacts"].sum()) > 10


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_golden(b2g, name):
    g = load(name)
    env, model, q, qd, tg, steps = make_golden.case_inputs(name)
    n = q.shape[0]
    batch = b2g.BatchBox2DWorld(n).loadWorld(model)
    batch.setWorldState(g["q0"], g["qd0"], reset_caches=True)
    for seg in range(tg.shape[0]):
        batch.setTargetVelocity(0, g["targets"][seg])
        batch.stepPhysics(int(g["steps"]))
        assert np.array_equal(batch.getWorldState()[0], g["traj_q"][:, seg]), "%s: segment %d" % (name, seg)
    gq, gqd = batch.getWorldState()
    assert np.array_equal(gq, g["q"]) and np.array_equal(gqd, g["qd"])
    assert np.array_equal(batch.bodies(), g["bodies"])
    counts, cio, _ = batch.contacts(max_per_world=80)
    assert np.array_equal(counts, g["n_contacts"])
    assert np.array_equal(cio[..., 2].sum(axis=1), g["n_touching"])
