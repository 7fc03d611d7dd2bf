"""Synthetic environment descriptions for the benchmark configs BASELINE.json names (host side, setup time).

`clutter_env` builds "shape C" (SURVEY.md §8): the articulated robot of test_env.yaml plus `n_objects` single-polygon
dynamic objects (convex, 3-8 vertices, circum-radius U(0.05, 0.15) m, mass U(0.05, 1) kg, contact friction
U(0.1, 1.5), ground friction U(0.05, 0.5), restitution 0 or 1).  The description is the plain-dict form that
`Environment.from_dict` (C ABI: b2g_env_create / b2g_env_add_*) consumes — the same structure the reference fills
from YAML into Box2DEnvironmentDescription (include/sim_env/Box2DIOUtils.h:18-70).

`clutter_states` draws per-world initial placements: every object gets its own cell of an 8 x 8 grid over the
3 m x 3 m region around the robot (cells under the robot are left empty), a jitter inside the cell, a random
heading and a random sliding velocity, so no two shapes overlap initially, every world of a batch is different,
and the clutter collides with itself and the robot within a 100-step rollout.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _robot_description(path):
    """Decode robots/*.yaml into the dict form (keys as in include/sim_env/Box2DIOUtils.h:94-303)."""
    import yaml
    with open(path) as f:
        node = yaml.safe_load(f)

    def pick(d, *names):
        for n in names:
            if n in d:
                return float(d[n])
        raise KeyError(names[0])

    links = [dict(name=str(l["name"]), geometry=[[float(v) for v in poly] for poly in l["geometry"]], mass=float(l["mass"]),
                  ground_friction=pick(l, "ground_friction", "trans_friction"),
                  ground_torque_integral=pick(l, "ground_torque_integral", "rot_friction"),
                  contact_friction=float(l["contact_friction"]), restitution=float(l["restitution"]))
             for l in node.get("links") or []]
    joints = [dict(name=str(j["name"]), link_a=str(j["link_a"]), link_b=str(j["link_b"]), axis=[float(v) for v in j["axis"]],
                   axis_orientation=float(j["axis_orientation"]), position_limits=[float(v) for v in j["position_limits"]],
                   velocity_limits=[float(v) for v in j["velocity_limits"]],
                   acceleration_limits=[float(v) for v in j["acceleration_limits"]], joint_type=str(j["joint_type"]),
                   actuated=bool(j["actuated"])) for j in node.get("joints") or []]
    ba = node["base_actuation"]
    return dict(name=str(node["name"]), static=False, links=links, joints=joints,
                base_actuation=dict(translational_velocity_limits=[float(v) for v in ba["translational_velocity_limits"]],
                                    translational_acceleration_limits=[float(v) for v in ba["translational_acceleration_limits"]],
                                    rotational_velocity_limits=[float(v) for v in ba["rotational_velocity_limits"]],
                                    rotational_acceleration_limits=[float(v) for v in ba["rotational_acceleration_limits"]],
                                    use_center_of_mass=bool(ba["use_center_of_mass"])))


def clutter_env(n_objects=32, seed=42, robot_yaml=None):
    rng = np.random.Generator(np.random.Philox(key=seed))
    robot = _robot_description(robot_yaml or os.path.join(_HERE, "data", "robots", "test_robot.yaml"))
    robot["name"] = "my_robot"
    env = dict(scale=10.0, world_bounds=[-20.0, -20.0, 20.0, 20.0], robots=[robot], objects=[], states={})
    for i in range(n_objects):
        nv = int(rng.integers(3, 9))
        radius = float(rng.uniform(0.05, 0.15))
        # convex by construction: vertices on a circle at sorted, well separated angles
        gaps = rng.uniform(0.6, 1.4, nv)
        ang = np.cumsum(gaps) * (2.0 * np.pi / gaps.sum()) + rng.uniform(0, 2.0 * np.pi)
        xy = np.stack([radius * np.cos(ang), radius * np.sin(ang)], axis=1).astype(np.float32)
        link = dict(name="link", geometry=[[float(v) for v in xy.reshape(-1)]], mass=float(np.float32(rng.uniform(0.05, 1.0))),
                    ground_friction=float(np.float32(rng.uniform(0.05, 0.5))), ground_torque_integral=float(np.float32(rng.uniform(0.01, 0.1))),
                    contact_friction=float(np.float32(rng.uniform(0.1, 1.5))), restitution=float(rng.integers(0, 2)))
        env["objects"].append(dict(name="clutter_%02d" % i, static=False, links=[link], joints=[], base_actuation=None))
    return env


def clutter_states(n_worlds, n_objects=32, seed=1, robot_dofs=7, speed=2.0, spin=3.0):
    """q, qd [n_worlds, robot_dofs + 3 * n_objects]: robot at the origin at rest, objects scattered (see module doc)
    and sliding with velocities U(-speed, speed) m/s, U(-spin, spin) rad/s so that the clutter collides with itself and
    with the robot inside a 100-step rollout."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    cells = []
    g, cell = 8, 3.0 / 8
    for iy in range(g):
        for ix in range(g):
            cx, cy = -1.5 + (ix + 0.5) * cell, -1.5 + (iy + 0.5) * cell
            # keep clear of the robot at the origin: its pose is the base COM, so it spans about x in [-0.45, 0.45],
            # y in [-0.5, 0.7]; an object reaches 0.18 m from its cell centre
            if -0.75 < cx < 0.75 and -0.8 < cy < 1.0:
                continue
            cells.append((cx, cy))
    cells = np.array(cells, np.float32)
    if len(cells) < n_objects:
        raise ValueError("at most %d objects fit the placement grid" % len(cells))
    nq = robot_dofs + 3 * n_objects
    q = np.zeros((n_worlds, nq), np.float32)
    qd = np.zeros((n_worlds, nq), np.float32)
    order = np.argsort(rng.random((n_worlds, len(cells))), axis=1)[:, :n_objects]
    centre = cells[order]
    jitter = rng.uniform(-0.03, 0.03, (n_worlds, n_objects, 2)).astype(np.float32)
    q[:, robot_dofs:].reshape(n_worlds, n_objects, 3)[:, :, 0:2] = centre + jitter
    q[:, robot_dofs:].reshape(n_worlds, n_objects, 3)[:, :, 2] = rng.uniform(-np.pi, np.pi, (n_worlds, n_objects)).astype(np.float32)
    v = qd[:, robot_dofs:].reshape(n_worlds, n_objects, 3)
    v[:, :, 0:2] = rng.uniform(-speed, speed, (n_worlds, n_objects, 2)).astype(np.float32)
    v[:, :, 2] = rng.uniform(-spin, spin, (n_worlds, n_objects)).astype(np.float32)
    return q, qd
