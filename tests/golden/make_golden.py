"""Generate the frozen trajectories under tests/golden/ FROM THE ORACLE (oracle/, Box2D 2.3.1 semantics as restated).

These are regression pins, not reference outputs: the reference (JoshuaHaustein/box2d_sim_env) cannot be built in this
environment and its own tests hold no numeric vectors, so parity stays "unpinned" at the Box2D boundary.
What the fixtures do guarantee: the oracle and the CUDA path keep producing, bit for bit, the trajectories that were
reviewed when the fixtures were frozen — on the GPU box the CUDA path is checked against them without the oracle.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (environment, worlds, segments x steps, seed)
    "shape_a": ("test_env.yaml", 16, (8, 10), 101),
    "shape_b": ("test_env_rigid_robot.yaml", 16, (8, 10), 202),
    "shape_c": ("clutter", 6, (5, 10), 303),
}


def case_inputs(name):
    """Seeded inputs of a case; shared with tests/test_golden.py so the GPU test needs nothing but the .npz."""
    import box2d_sim_env_b200 as b2g
    from box2d_sim_env_b200 import envgen
    from parity_utils import random_states, random_targets
    env_name, n, (segments, steps), seed = CASES[name]
    if env_name == "clutter":
        env = envgen.clutter_env()
        model = b2g.Model(env)
        q, qd = envgen.clutter_states(n, seed=seed, speed=3.0)
    else:
        env = b2g.data_path(env_name)
        model = b2g.Model(env)
        q, qd = random_states(model, n, seed=seed, spread=0.6)
    tg = random_targets(model, 0, n, seed=seed + 1, segments=segments)
    return env, model, q, qd, tg, steps


def run_oracle(name):
    from oracle import b2o
    from box2d_sim_env_b200 import envgen
    env, model, q, qd, tg, steps = case_inputs(name)
    env_dict = envgen.clutter_env() if CASES[name][0] == "clutter" else b2o.load_env_yaml(env)
    oenv = b2o.OracleEnv(env_dict)
    n = q.shape[0]
    out = dict(q0=q, qd0=qd, targets=tg, steps=np.int32(steps))
    fq, fqd, bodies, ncont, touching = [], [], [], [], []
    for w in range(n):
        ow = b2o.OracleWorld(oenv)
        ow.set_state(q[w], qd[w])
        traj_q = []
        for seg in range(tg.shape[0]):
            ow.set_target_velocity(0, tg[seg][w])
            ow.step(steps)
            traj_q.append(ow.get_state()[0])
        a, b = ow.get_state()
        fq.append(a)
        fqd.append(b)
        bodies.append(ow.bodies())
        io, _ = ow.contacts()
        ncont.append(len(io))
        touching.append(int(io[:, 2].sum()) if len(io) else 0)
        out.setdefault("traj_q", []).append(np.stack(traj_q))
    out.update(q=np.stack(fq), qd=np.stack(fqd), bodies=np.stack(bodies), n_contacts=np.array(ncont, np.int32),
               n_touching=np.array(touching, np.int32), traj_q=np.stack(out["traj_q"]))
    return out


if __name__ == "__main__":
    for name in CASES:
        data = run_oracle(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print("%s: %d worlds, %d live contacts, %d touching -> %s (%d bytes)" %
              (name, data["q"].shape[0], int(data["n_contacts"].sum()), int(data["n_touching"].sum()), path, os.path.getsize(path)))
