"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).

ctypes driver for ``oracle/libb2o.so`` (the CPU restatement of Box2DWorld + Box2D 2.3.1).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The YAML reading here uses PyYAML and is deliberately independent of the product's
own C++ loader (box2d_sim_env_b200/csrc/yaml_env.cpp) so that the two check each other.

Reference behaviour followed: /root/reference/src/sim_env/Box2DIOUtils.cpp:16-92 (file resolution and
environment decode), /root/reference/include/sim_env/Box2DIOUtils.h:94-303 (link / joint / robot decode).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
# "" = the oracle proper (deterministic sin/cos shared with the CUDA path); "_libm" = the sensitivity variant whose b2Rot
# calls libm sinf / cosf the way upstream Box2D does (oracle/Makefile: libb2o_libm.so, -DB2O_LIBM_TRIG)
_VARIANT = ""


def build(force=False):
    target = "libb2o%s.so" % _VARIANT
    so = os.path.join(_HERE, target)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, target], stdout=subprocess.DEVNULL)
    return so


def load_variant(suffix):
    """A second, independent copy of this module bound to libb2o<suffix>.so (e.g. "_libm"): every class of the copy
    calls its own lib(), so both oracles can live in one process."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle.b2o" + suffix, os.path.abspath(__file__))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod._VARIANT = suffix
    return mod


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.b2o_last_error.restype = C.c_char_p
        for name in ("b2o_env_new", "b2o_world_new", "b2o_batch_new", "b2o_batch_world"):
            getattr(L, name).restype = C.c_void_p
        L.b2o_env_new.argtypes = [C.c_float, C.c_void_p]
        L.b2o_env_free.argtypes = [C.c_void_p]
        L.b2o_env_add_object.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.b2o_env_add_link.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p] + [C.c_float] * 5
        L.b2o_env_add_polygon.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.b2o_env_add_ball.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float]
        L.b2o_world_set_object_dofs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_get_object_dofs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_ball_approximation.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_env_add_joint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                        C.c_void_p, C.c_int, C.c_char_p]
        L.b2o_env_add_state.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_world_new.argtypes = [C.c_void_p]
        L.b2o_world_free.argtypes = [C.c_void_p]
        L.b2o_world_info.argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_world_set_params.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.b2o_world_set_continuous.argtypes = [C.c_void_p, C.c_int]
        L.b2o_world_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.b2o_world_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2o_world_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2o_world_set_object_pose.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.b2o_world_get_object_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_set_control_efforts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b2o_world_set_target_velocity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b2o_world_set_target_position.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float]
        L.b2o_world_object_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        for name in ("b2o_world_get_bodies", "b2o_world_get_body_model", "b2o_world_get_fat_aabbs"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_world_get_fixture_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2o_world_get_joints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2o_world_get_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_world_check_collision.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_world_objects_in_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_is_physically_feasible.argtypes = [C.c_void_p]
        L.b2o_world_step_record.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_world_set_link_enabled.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_world_get_link_states.argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_world_get_link_aabb.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_get_object_aabb.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_world_set_link_property.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.b2o_world_get_link_property.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_world_get_link_property.restype = C.c_float
        L.b2o_world_trace_islands.argtypes = [C.c_void_p, C.c_int]
        L.b2o_world_get_island_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_world_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_world_controller_efforts.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_sincos.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_polygon_set.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.b2o_collide_polygons.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.b2o_distance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_time_of_impact.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.b2o_batch_new.argtypes = [C.c_void_p, C.c_int]
        L.b2o_batch_free.argtypes = [C.c_void_p]
        L.b2o_batch_world.argtypes = [C.c_void_p, C.c_int]
        L.b2o_batch_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------------------------------------------
# environment description (plain dicts) from YAML — independent of the product loader
# ---------------------------------------------------------------------------------------------------
def _decode_link(node):
    # key aliases: the reference data uses trans_friction / rot_friction, its decoder reads
    # ground_friction / ground_torque_integral (Box2DIOUtils.h:133-134); see SURVEY.md Appendix C-1.
    gf = node["ground_friction"] if "ground_friction" in node else node["trans_friction"]
    gti = node["ground_torque_integral"] if "ground_torque_integral" in node else node["rot_friction"]
    return dict(name=str(node["name"]), geometry=[[float(v) for v in poly] for poly in node["geometry"]],
                mass=float(node["mass"]), ground_friction=float(gf), ground_torque_integral=float(gti),
                contact_friction=float(node["contact_friction"]), restitution=float(node["restitution"]),
                # yaml-cpp pairs [[x, y, z], radius] (Box2DIOUtils.h:139-142)
                balls=[[float(c[0]), float(c[1]), float(c[2]), float(r)] for c, r in (node.get("balls") or [])])


def _decode_joint(node):
    return dict(name=str(node["name"]), link_a=str(node["link_a"]), link_b=str(node["link_b"]),
                axis=[float(v) for v in node["axis"]], axis_orientation=float(node["axis_orientation"]),
                position_limits=[float(v) for v in node["position_limits"]],
                velocity_limits=[float(v) for v in node["velocity_limits"]],
                acceleration_limits=[float(v) for v in node["acceleration_limits"]],
                joint_type=str(node["joint_type"]), actuated=bool(node["actuated"]))


def _decode_object(path, name, static, robot):
    import yaml
    with open(path) as f:
        node = yaml.safe_load(f)
    obj = dict(name=name or str(node["name"]), static=static,
               links=[_decode_link(l) for l in (node.get("links") or [])],
               joints=[_decode_joint(j) for j in (node.get("joints") or [])], base_actuation=None)
    if robot:
        ba = node.get("base_actuation")
        if ba is not None:
            obj["static"] = False
            obj["base_actuation"] = dict(
                translational_velocity_limits=[float(v) for v in ba["translational_velocity_limits"]],
                translational_acceleration_limits=[float(v) for v in ba["translational_acceleration_limits"]],
                rotational_velocity_limits=[float(v) for v in ba["rotational_velocity_limits"]],
                rotational_acceleration_limits=[float(v) for v in ba["rotational_acceleration_limits"]],
                use_center_of_mass=bool(ba["use_center_of_mass"]))
        else:
            obj["static"] = True
    return obj


def load_env_yaml(path):
    import yaml
    root = os.path.dirname(os.path.abspath(path))
    with open(path) as f:
        node = yaml.safe_load(f)
    env = dict(scale=float(node.get("scale", 1.0)),
               world_bounds=[float(v) for v in node.get("world_bounds", [0, 0, 0, 0])],
               robots=[], objects=[], states={})
    for e in node.get("robots") or []:
        env["robots"].append(_decode_object(os.path.join(root, e["filename"]), e.get("name"), False, True))
    for e in node.get("objects") or []:
        env["objects"].append(_decode_object(os.path.join(root, e["filename"]), e.get("name"),
                                             bool(e.get("static", False)), False))
    for s in node.get("states") or []:
        env["states"][str(s["name"])] = dict(configuration=[float(v) for v in s["state"]["configuration"]],
                                             velocity=[float(v) for v in s["state"]["velocity"]])
    return env


class OracleEnv:
    def __init__(self, env):
        L = lib()
        self.desc = env
        b = _f32(env.get("world_bounds", [0, 0, 0, 0]))
        self.h = C.c_void_p(L.b2o_env_new(float(env["scale"]), _ptr(b)))
        for is_robot, key in ((1, "robots"), (0, "objects")):
            for obj in env[key]:
                ba = obj.get("base_actuation")
                lim = None
                use_com = 1
                if ba:
                    lim = _f32(list(ba["translational_velocity_limits"]) + list(ba["translational_acceleration_limits"]) +
                               list(ba["rotational_velocity_limits"]) + list(ba["rotational_acceleration_limits"]))
                    use_com = int(ba.get("use_center_of_mass", True))
                oi = L.b2o_env_add_object(self.h, obj["name"].encode(), is_robot, int(obj["static"]), _ptr(lim), use_com)
                for link in obj["links"]:
                    li = L.b2o_env_add_link(self.h, is_robot, oi, link["name"].encode(), link["mass"],
                                            link["ground_friction"], link["ground_torque_integral"],
                                            link["contact_friction"], link["restitution"])
                    for poly in link["geometry"]:
                        p = _f32(poly)
                        L.b2o_env_add_polygon(self.h, is_robot, oi, li, _ptr(p), len(p))
                    for ball in link.get("balls") or []:
                        c = _f32(ball[:3])
                        L.b2o_env_add_ball(self.h, is_robot, oi, li, _ptr(c), float(ball[3]))
                for j in obj["joints"]:
                    params = _f32(list(j["axis"]) + list(j["position_limits"]) + list(j["velocity_limits"]) +
                                  list(j["acceleration_limits"]) + [j["axis_orientation"]])
                    L.b2o_env_add_joint(self.h, is_robot, oi, j["name"].encode(), j["link_a"].encode(),
                                        j["link_b"].encode(), _ptr(params), int(j["actuated"]), j["joint_type"].encode())
        for name, st in env.get("states", {}).items():
            c = _f32(st["configuration"])
            v = _f32(st["velocity"])
            L.b2o_env_add_state(self.h, name.encode(), _ptr(c), _ptr(v), len(c))

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2o_env_free(self.h)
            self.h = None


class OracleWorld:
    """One CPU world with the reference's Box2DWorld semantics."""

    MAXC = 256

    def __init__(self, env, _handle=None):
        L = lib()
        self.env = env if isinstance(env, OracleEnv) else OracleEnv(env)
        self._owned = _handle is None
        self.h = C.c_void_p(L.b2o_world_new(self.env.h)) if _handle is None else C.c_void_p(_handle)
        if not self.h:
            raise RuntimeError(L.b2o_last_error().decode())
        info = np.zeros(6, np.int32)
        L.b2o_world_info(self.h, _ptr(info))
        self.n_bodies, self.n_proxies, self.n_joints, self.nq, self.n_objects = [int(x) for x in info[:5]]
        self.objects = []
        for i in range(self.n_objects):
            oi = np.zeros(6, np.int32)
            nm = C.create_string_buffer(128)
            L.b2o_world_object_info(self.h, i, _ptr(oi), nm, 128)
            self.objects.append(dict(name=nm.value.decode(), is_robot=bool(oi[0]), is_static=bool(oi[1]),
                                     n_dofs=int(oi[2]), n_links=int(oi[3]), n_joints=int(oi[4]), first_body=int(oi[5])))

    def __del__(self):
        if getattr(self, "h", None) and self._owned:
            lib().b2o_world_free(self.h)
            self.h = None

    def object_index(self, name):
        for i, o in enumerate(self.objects):
            if o["name"] == name:
                return i
        raise KeyError(name)

    def set_params(self, dt=0.01, vel_iters=15, pos_iters=15):
        lib().b2o_world_set_params(self.h, dt, vel_iters, pos_iters)

    def set_continuous(self, flag):
        lib().b2o_world_set_continuous(self.h, int(flag))

    def step(self, steps=1, execute_controllers=True, allow_sleeping=True):
        if lib().b2o_world_step(self.h, steps, int(execute_controllers), int(allow_sleeping)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def get_state(self):
        q = np.zeros(self.nq, np.float32)
        qd = np.zeros(self.nq, np.float32)
        lib().b2o_world_get_state(self.h, _ptr(q), _ptr(qd))
        return q, qd

    def set_state(self, q, qd):
        q = _f32(q)
        qd = _f32(qd)
        assert q.size == self.nq and qd.size == self.nq
        if lib().b2o_world_set_state(self.h, _ptr(q), _ptr(qd)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def set_object_pose(self, obj, x, y, theta):
        lib().b2o_world_set_object_pose(self.h, obj, x, y, theta)

    def get_object_pose(self, obj):
        p = np.zeros(3, np.float32)
        lib().b2o_world_get_object_pose(self.h, obj, _ptr(p))
        return p

    def set_object_dofs(self, obj, values, indices=None, velocities=False):
        """Box2DObject::setDOFPositions / setDOFVelocities(values, indices); indices None = all DOFs."""
        v = _f32(values)
        ix = np.ascontiguousarray(indices, np.int32) if indices is not None else None
        if lib().b2o_world_set_object_dofs(self.h, obj, int(velocities), _ptr(ix), 0 if ix is None else len(ix), _ptr(v)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def get_object_dofs(self, obj, indices=None, velocities=False):
        ix = np.ascontiguousarray(indices, np.int32) if indices is not None else None
        out = np.zeros(self.objects[obj]["n_dofs"] if ix is None else len(ix), np.float32)
        lib().b2o_world_get_object_dofs(self.h, obj, int(velocities), _ptr(ix), 0 if ix is None else len(ix), _ptr(out))
        return out

    def ball_approximation(self, obj):
        """Box2DObject::getBallApproximation: [n_balls, 4] = centre (world frame, user units), radius."""
        n = lib().b2o_world_ball_approximation(self.h, obj, None)
        out = np.zeros((n, 4), np.float32)
        if n:
            lib().b2o_world_ball_approximation(self.h, obj, _ptr(out))
        return out

    def set_target_velocity(self, obj, v):
        v = _f32(v)
        if lib().b2o_world_set_target_velocity(self.h, obj, _ptr(v), len(v)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def set_target_position(self, obj, q, kp):
        q = _f32(q)
        if lib().b2o_world_set_target_position(self.h, obj, _ptr(q), len(q), C.c_float(kp)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def set_control_efforts(self, obj, efforts):
        """Install a controller that returns these efforts every step (Box2DRobot::setController + commandEfforts)."""
        e = _f32(efforts)
        if lib().b2o_world_set_control_efforts(self.h, obj, _ptr(e), len(e)):
            raise RuntimeError(lib().b2o_last_error().decode())

    def controller_efforts(self, obj):
        out = np.zeros(64, np.float32)
        n = lib().b2o_world_controller_efforts(self.h, obj, _ptr(out))
        return out[:n].copy()

    def link_aabb(self, body_index):
        out = np.zeros(4, np.float32)
        lib().b2o_world_get_link_aabb(self.h, body_index, _ptr(out))
        return out

    def object_aabb(self, obj):
        out = np.zeros(4, np.float32)
        lib().b2o_world_get_object_aabb(self.h, obj, _ptr(out))
        return out

    def link_states(self):
        """[n_bodies, 8]: Box2DLink::getPose(3), getVelocityVector(3), getCenterOfMass(2); row 0 (ground) is 0"""
        out = np.zeros((self.n_bodies, 8), np.float32)
        lib().b2o_world_get_link_states(self.h, _ptr(out))
        return out

    def bodies(self):
        """[n_bodies, 16]: xf.p(2) xf.q.s xf.q.c c(2) a c0(2) a0 v(2) w sleepTime awake type"""
        out = np.zeros((self.n_bodies, 16), np.float32)
        lib().b2o_world_get_bodies(self.h, _ptr(out))
        return out

    def body_model(self):
        out = np.zeros((self.n_bodies, 8), np.float32)
        lib().b2o_world_get_body_model(self.h, _ptr(out))
        return out

    def fixture_model(self):
        io = np.zeros((self.n_proxies, 3), np.int32)
        fo = np.zeros((self.n_proxies, 38), np.float32)
        lib().b2o_world_get_fixture_model(self.h, _ptr(io), _ptr(fo))
        return io, fo

    def fat_aabbs(self):
        out = np.zeros((self.n_proxies, 4), np.float32)
        lib().b2o_world_get_fat_aabbs(self.h, _ptr(out))
        return out

    def joints(self):
        io = np.zeros((self.n_joints, 6), np.int32)
        fo = np.zeros((self.n_joints, 15), np.float32)
        lib().b2o_world_get_joints(self.h, _ptr(io), _ptr(fo))
        return io, fo

    def contacts(self):
        io = np.zeros((self.MAXC, 8), np.int32)
        fo = np.zeros((self.MAXC, 12), np.float32)
        n = lib().b2o_world_get_contacts(self.h, _ptr(io), _ptr(fo), self.MAXC)
        return io[:n].copy(), fo[:n].copy()

    def check_collision(self):
        io = np.zeros((self.MAXC, 4), np.int32)
        fo = np.zeros((self.MAXC, 6), np.float32)
        n = lib().b2o_world_check_collision(self.h, _ptr(io), _ptr(fo), self.MAXC)
        return io[:n].copy(), fo[:n].copy()

    def objects_in_aabb(self, aabb, exclude_robots=True, n_objects=None):
        n = n_objects if n_objects is not None else len(self.env.desc["robots"]) + len(self.env.desc["objects"])
        hits = np.zeros(n, np.int32)
        box = _f32(aabb)
        lib().b2o_world_objects_in_aabb(self.h, _ptr(box), int(exclude_robots), _ptr(hits))
        return hits.astype(bool)

    def is_physically_feasible(self):
        return bool(lib().b2o_world_is_physically_feasible(self.h))

    def step_record(self, steps=1, execute_controllers=True, allow_sleeping=True):
        io = np.zeros((self.MAXC, 4), np.int32)
        fo = np.zeros((self.MAXC, 6), np.float32)
        n = lib().b2o_world_step_record(self.h, steps, int(execute_controllers), int(allow_sleeping), _ptr(io), _ptr(fo),
                                        self.MAXC)
        return io[:n].copy(), fo[:n].copy()

    def set_link_enabled(self, body_index, enabled):
        lib().b2o_world_set_link_enabled(self.h, body_index, int(enabled))

    # Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction and getters
    LINK_PROPERTIES = dict(mass=0, ground_friction=1, ground_torque_integral=2, contact_friction=3, inertia=4,
                           com_inertia=5, ground_friction_limit_force=6, ground_friction_limit_torque=7)

    def set_link_property(self, body_index, prop, value):
        lib().b2o_world_set_link_property(self.h, body_index, self.LINK_PROPERTIES[prop], float(value))

    def get_link_property(self, body_index, prop):
        return float(lib().b2o_world_get_link_property(self.h, body_index, self.LINK_PROPERTIES[prop]))

    def trace_islands(self, flag=True):
        lib().b2o_world_trace_islands(self.h, int(flag))

    def island_trace(self):
        buf = np.zeros(4096, np.int32)
        n = lib().b2o_world_get_island_trace(self.h, _ptr(buf), 4096)
        v = buf[:n].tolist()
        out = []
        p = 1
        for _ in range(v[0] if v else 0):
            nb = v[p]; bodies = v[p + 1:p + 1 + nb]; p += 1 + nb
            nj = v[p]; joints = v[p + 1:p + 1 + nj]; p += 1 + nj
            nc = v[p]; cs = [(v[p + 1 + 2 * k], v[p + 2 + 2 * k]) for k in range(nc)]; p += 1 + 2 * nc
            out.append(dict(bodies=bodies, joints=joints, contacts=cs))
        return out

    def stats(self):
        s = np.zeros(3, np.int32)
        lib().b2o_world_stats(self.h, _ptr(s))
        return dict(islands=int(s[0]), toi_events=int(s[1]), toi_calls=int(s[2]))


class OracleBatch:
    """N independent CPU worlds stepped by a std::thread pool (the CPU baseline of SURVEY.md §8d)."""

    def __init__(self, env, n):
        self.env = env if isinstance(env, OracleEnv) else OracleEnv(env)
        self.n = n
        self.h = C.c_void_p(lib().b2o_batch_new(self.env.h, n))
        if not self.h:
            raise RuntimeError(lib().b2o_last_error().decode())
        self.world0 = OracleWorld(self.env, _handle=lib().b2o_batch_world(self.h, 0))
        self.nq = self.world0.nq

    def world(self, i):
        return OracleWorld(self.env, _handle=lib().b2o_batch_world(self.h, i))

    def rollout(self, q0, qd0, robot_idx, targets, steps_per_target, n_threads):
        """targets: [T, n, ndof] or None.  Returns (q, qd) after T*steps_per_target steps."""
        q0 = _f32(q0) if q0 is not None else None
        qd0 = _f32(qd0) if qd0 is not None else None
        if targets is not None:
            targets = _f32(targets)
            T, n, ndof = targets.shape
            assert n == self.n
        else:
            T, ndof = 1, 0
        q = np.zeros((self.n, self.nq), np.float32)
        qd = np.zeros((self.n, self.nq), np.float32)
        failed = lib().b2o_batch_rollout(self.h, _ptr(q0), _ptr(qd0), robot_idx, _ptr(targets), ndof, T,
                                         steps_per_target, n_threads, _ptr(q), _ptr(qd))
        if failed:
            raise RuntimeError("%d oracle worlds failed" % failed)
        return q, qd

    def __del__(self):
        if getattr(self, "h", None):
            self.world0 = None
            lib().b2o_batch_free(self.h)
            self.h = None


def sincos(x):
    x = _f32(x)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    lib().b2o_sincos(_ptr(x), _ptr(s), _ptr(c), x.size)
    return s, c


def polygon_set(points, density=1.0):
    p = _f32(points).reshape(-1)
    out = np.zeros(38, np.float32)
    n = lib().b2o_polygon_set(_ptr(p), len(p) // 2, density, _ptr(out))
    return dict(count=n, vertices=out[:16].reshape(8, 2)[:n].copy(), normals=out[16:32].reshape(8, 2)[:n].copy(),
                centroid=out[32:34].copy(), mass=float(out[34]), center=out[35:37].copy(), I=float(out[37]))


def distance(polyA, poseA, polyB, poseB, use_radii=False):
    """b2Distance between two polygons: dict(distance, iterations, point_a, point_b)."""
    a, b = _f32(polyA).reshape(-1), _f32(polyB).reshape(-1)
    pa, pb = _f32(poseA), _f32(poseB)
    out = np.zeros(6, np.float32)
    lib().b2o_distance(_ptr(a), len(a) // 2, _ptr(pa), _ptr(b), len(b) // 2, _ptr(pb), int(use_radii), _ptr(out))
    return dict(distance=float(out[0]), iterations=int(out[1]), point_a=out[2:4].copy(), point_b=out[4:6].copy())


def time_of_impact(polyA, poseA0, poseA1, polyB, poseB0, poseB1):
    """b2TimeOfImpact over [0, 1] for two polygons sweeping between two poses: (state, t); state 3 = touching,
    4 = separated, 2 = overlapped, 1 = failed."""
    a, b = _f32(polyA).reshape(-1), _f32(polyB).reshape(-1)
    out = np.zeros(2, np.float32)
    lib().b2o_time_of_impact(_ptr(a), len(a) // 2, _ptr(_f32(poseA0)), _ptr(_f32(poseA1)), _ptr(b), len(b) // 2,
                             _ptr(_f32(poseB0)), _ptr(_f32(poseB1)), _ptr(out))
    return int(out[0]), float(out[1])


def collide_polygons(polyA, poseA, polyB, poseB):
    a = _f32(polyA).reshape(-1)
    b = _f32(polyB).reshape(-1)
    pa = _f32(poseA)
    pb = _f32(poseB)
    io = np.zeros(4, np.int32)
    fo = np.zeros(8, np.float32)
    lib().b2o_collide_polygons(_ptr(a), len(a) // 2, _ptr(pa), _ptr(b), len(b) // 2, _ptr(pb), _ptr(io), _ptr(fo))
    return dict(point_count=int(io[0]), type=int(io[1]), ids=[int(io[2]) & 0xFFFFFFFF, int(io[3]) & 0xFFFFFFFF],
                local_normal=fo[0:2].copy(), local_point=fo[2:4].copy(), points=fo[4:8].reshape(2, 2).copy())
