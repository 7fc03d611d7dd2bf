"""Helpers shared by the parity tests: scenario generation and oracle <-> CUDA comparison."""
import numpy as np

import box2d_sim_env_b200 as b2g
from oracle import b2o

ENV_A = "test_env.yaml"
ENV_B = "test_env_rigid_robot.yaml"


def load_env_dict(name):
    return b2o.load_env_yaml(b2g.data_path(name))


ENV_2 = "variant:env2"   # the reference's test_data/test_env2.yaml: base-only robot, obj2 (two hinged bars) DYNAMIC
ENV_3 = "variant:env3"   # the reference's test_data/test_env3.yaml: one-link frictionless fingers, obj2 dynamic


def env_variant(name):
    """The reference ships two more environments (test_data/test_env2.yaml + robots/test_robot2.yaml, test_env3.yaml +
    robots/test_robot3.yaml).  They are the shipped shape-A files with parts removed / parameters changed, so they are
    rebuilt here from this repo's shape-A description; tests/test_host.py checks the result against the reference's own
    files where /root/reference exists."""
    env = load_env_dict(ENV_A)
    robot = env["robots"][0]
    ba = robot["base_actuation"]
    ba["translational_velocity_limits"] = [-10.0, 10.0]
    ba["rotational_velocity_limits"] = [-3.14, 3.14]
    if name == ENV_2:
        ba["translational_acceleration_limits"] = [-1.8, 1.8]
        robot["links"] = robot["links"][:1]
        robot["links"][0]["ground_torque_integral"] = 0.0
        robot["joints"] = []
        ndof = 3
    elif name == ENV_3:
        keep = ("test_robot_base", "test_robot_finger_l_1", "test_robot_finger_r_1")
        robot["links"] = [l for l in robot["links"] if l["name"] in keep]
        for l in robot["links"][1:]:
            l["ground_friction"] = 0.0
            l["ground_torque_integral"] = 0.0
        robot["joints"] = [j for j in robot["joints"] if j["name"] in ("finger_l_joint_1", "finger_r_joint_1")]
        ndof = 5
    else:
        raise KeyError(name)
    env["world_bounds"] = [0.0, 0.0, 0.0, 0.0]          # no world_bounds key: defaults (Box2DIOUtils.cpp:79-84)
    env["objects"][1]["static"] = False                  # obj2 dynamic: 3 base DOFs + its hinge
    env["states"]["my_robot"] = dict(configuration=[0.0] * ndof, velocity=[0.0] * ndof)
    env["states"]["obj2"] = dict(configuration=[2.9, 1.2, -6.12, 0.0], velocity=[0.0, 0.0, 0.0, -0.5])
    return env


def env_dict(name):
    return env_variant(name) if name.startswith("variant:") else load_env_dict(name)


def random_states(model, n, seed, spread=1.2, vel=1.0):
    """Random world states inside the DOF limits: objects scattered around the origin so contacts happen."""
    rng = np.random.default_rng(seed)
    lim = model.dof_limits()
    q = np.zeros((n, model.nq), np.float32)
    qd = np.zeros((n, model.nq), np.float32)
    for o in model.objects:
        off, nd = o["dof_offset"], o["n_dofs"]
        base = 0 if o["is_static"] else 3
        if base:
            q[:, off:off + 2] = rng.uniform(-spread, spread, (n, 2))
            q[:, off + 2] = rng.uniform(-np.pi, np.pi, n)
            qd[:, off:off + 2] = rng.uniform(-vel, vel, (n, 2))
            qd[:, off + 2] = rng.uniform(-vel, vel, n)
        for k in range(base, nd):
            lo, hi = lim[off + k, 0], lim[off + k, 1]
            q[:, off + k] = rng.uniform(0.9 * lo, 0.9 * hi, n)
            qd[:, off + k] = rng.uniform(-vel, vel, n)
    return q, qd


def base_front_edge(fat_aabbs_world0, scale=10.0):
    """y of the robot base rectangle's front (upper) edge in the robot frame, read off the fat AABB of fixture 1 of a freshly
    loaded world (robot at the origin): fat = tight + polygon radius (0.01) + aabb extension (0.1), Box2D units."""
    return (float(fat_aabbs_world0[1][3]) - 0.11) / scale


def feasible_states(model, n, seed, ytop):
    """Physically feasible (no initial interpenetration) contact-rich states for the shipped env shapes and their variants:
    the robot at the origin at rest, obj1 in the robot's palm -- a quarter of the worlds with the box resting against the
    base's front edge inside the polygon skin (touching from the first step) --, obj2 where its YAML puts it.  The same
    placement bench.py uses for BASELINE.json configs[1] / [2]."""
    rng = np.random.default_rng(seed)
    q = np.zeros((n, model.nq), np.float32)
    qd = np.zeros((n, model.nq), np.float32)
    for o in model.objects:
        off = o["dof_offset"]
        if o["name"] == "obj1":
            resting = rng.uniform(0, 1, n) < 0.25
            cx = rng.uniform(-0.17, 0.17, n)
            cy = np.where(resting, ytop + rng.uniform(0.0002, 0.0015, n) + 0.05, rng.uniform(ytop + 0.08, ytop + 0.30, n))
            th = np.where(resting, 0.0, rng.uniform(-np.pi, np.pi, n))
            c, s = np.cos(th), np.sin(th)
            q[:, off] = cx - (c * 0.05 - s * 0.05)
            q[:, off + 1] = cy - (s * 0.05 + c * 0.05)
            q[:, off + 2] = th
        elif o["name"] == "obj2" and not o["is_static"]:
            q[:, off:off + 3] = [2.9, 1.2, -6.12]
    return q, qd


def random_targets(model, robot, n, seed, segments=1):
    rng = np.random.default_rng(seed)
    lim = model.dof_limits()
    o = model.objects[robot]
    off, nd = o["dof_offset"], o["n_dofs"]
    t = rng.uniform(lim[off:off + nd, 2], lim[off:off + nd, 3], (segments, n, nd)).astype(np.float32)
    return t


def oracle_worlds(env_dict, n, continuous=True):
    env = b2o.OracleEnv(env_dict)
    ws = [b2o.OracleWorld(env) for _ in range(n)]
    for w in ws:
        w.set_continuous(continuous)
    return ws


def contact_table(io, fo=None):
    """Canonical (fixtureA, fixtureB, touching, pointCount) rows in creation order."""
    return [tuple(int(x) for x in r[[0, 1, 2, 4]]) for r in io]


def compare_world(batch_bodies, batch_contacts, w, oracle_world, atol=0.0):
    """Returns a list of human-readable differences between world `w` of the batch and the oracle world."""
    diffs = []
    ob = oracle_world.bodies()
    gb = batch_bodies[w]
    names = ["px", "py", "qs", "qc", "cx", "cy", "a", "c0x", "c0y", "a0", "vx", "vy", "w", "sleep", "awake", "type"]
    bad = np.argwhere(~((np.abs(ob - gb) <= atol) | (ob == gb) | (np.isnan(ob) & np.isnan(gb))))
    for b, f in bad:
        diffs.append("body %d %s: oracle %.9g cuda %.9g" % (b, names[f], ob[b, f], gb[b, f]))
    counts, cio, cfo = batch_contacts
    oio, ofo = oracle_world.contacts()
    n = int(counts[w])
    if n != len(oio):
        diffs.append("contact count: oracle %d cuda %d" % (len(oio), n))
    else:
        for k in range(n):
            if not np.array_equal(oio[k, :6], cio[w, k, :6]):
                diffs.append("contact %d ints: oracle %s cuda %s" % (k, oio[k].tolist(), cio[w, k].tolist()))
            elif oio[k, 4] > 0:
                npts = int(oio[k, 4])
                if not np.array_equal(oio[k, 6:6 + npts], cio[w, k, 6:6 + npts]):
                    diffs.append("contact %d ids: oracle %s cuda %s" % (k, oio[k, 6:].tolist(), cio[w, k, 6:].tolist()))
                sel = list(range(4 + 4 * npts))
                a, b = ofo[k, sel], cfo[w, k, sel]
                if not np.all((np.abs(a - b) <= atol) | (a == b)):
                    diffs.append("contact %d floats: oracle %s cuda %s" % (k, a.tolist(), b.tolist()))
    return diffs


def env_prismatic(actuated=False, static_cabinet=False, limits=(-0.1, 0.2)):
    """Shape A's robot in front of a drawer: a cabinet link and a slider link joined by a PRISMATIC joint along the cabinet's
    x axis (the reference constructs a b2PrismaticJoint for `joint_type: prismatic`, Box2DWorld.cpp:920-940, and can step and
    read it but not place it: no `states` entry for the drawer, it is created at the origin).  The robot starts turned by
    -90 degrees behind the slider, which sits in its palm, so that driving the base along +x pushes the slider along its
    axis into the upper limit and then pushes the whole drawer.  (`axis` is the joint anchor on the cabinet in BOX2D units: the
    reference does not scale it for prismatic joints, :923; 3.5 puts the slider 0.05 m inside its travel when the world is
    created.  Every link of a static object is a static body, so the drawer is dynamic.)"""
    env = load_env_dict(ENV_A)
    env["objects"] = [dict(
        name="drawer", static=bool(static_cabinet), base_actuation=None,
        links=[dict(name="cabinet", geometry=[[0.3, -0.2, 0.5, -0.2, 0.5, 0.2, 0.3, 0.2]], mass=5.0, ground_friction=0.3,
                    ground_torque_integral=0.2, contact_friction=0.8, restitution=0.2, balls=[]),
               dict(name="slider", geometry=[[-0.05, -0.05, 0.2, -0.05, 0.2, 0.05, -0.05, 0.05]], mass=0.2, ground_friction=0.05,
                    ground_torque_integral=0.01, contact_friction=0.8, restitution=0.1, balls=[])],
        joints=[dict(name="rail", link_a="cabinet", link_b="slider", axis=[3.5, 0.0], axis_orientation=0.0,
                     position_limits=[float(limits[0]), float(limits[1])], velocity_limits=[-5.0, 5.0],
                     acceleration_limits=[-0.6, 0.8], joint_type="prismatic", actuated=bool(actuated))])]
    env["states"] = {"my_robot": dict(configuration=[-0.5, 0.0, -1.5707964, 0.0, 0.0, 0.0, 0.0], velocity=[0.0] * 7)}
    return env
