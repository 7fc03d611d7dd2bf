"""ctypes mirror of include/b2g.h with the reference's method names.

Reference interface mirrored (file:line in /root/reference):
  Box2DWorld::loadWorld            include/sim_env/Box2DWorld.h:751
  stepPhysics(steps, ctrl, sleep)  :780            setPhysicsTimeStep/... :784-786
  getWorldState / setWorldState    :823-825        saveState/restoreState/dropState :820-822
  checkCollision(contacts)         :797            clone :748 (here: a second batch of the same model)
  Box2DRobotVelocityController::setTargetVelocity  include/sim_env/Box2DController.h:23
"""
import ctypes as C
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class B2GError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b2g error %d: %s" % (code, msg))
        self.code = code


def data_path(name="test_env.yaml"):
    return os.path.join(_HERE, "data", name)


class _ModelInfo(C.Structure):
    _fields_ = [("n_bodies", C.c_int32), ("n_fixtures", C.c_int32), ("n_joints", C.c_int32), ("n_objects", C.c_int32),
                ("n_robots", C.c_int32), ("nq", C.c_int32), ("scale", C.c_float)]


class _ObjectInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("is_robot", C.c_int32), ("is_static", C.c_int32), ("n_dofs", C.c_int32),
                ("dof_offset", C.c_int32), ("first_body", C.c_int32), ("n_links", C.c_int32), ("n_joints", C.c_int32),
                ("n_balls", C.c_int32)]


class _LinkInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("object", C.c_int32), ("is_base_link", C.c_int32), ("is_static", C.c_int32),
                ("parent_joint", C.c_int32), ("n_fixtures", C.c_int32)]


class _JointInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("kind", C.c_int32), ("object", C.c_int32), ("body_a", C.c_int32),
                ("body_b", C.c_int32), ("dof_index", C.c_int32), ("actuated", C.c_int32), ("limited", C.c_int32)]


# every symbol include/b2g.h declares (tests/test_abi.py checks the library exports them all)
ABI_SYMBOLS = [
    "b2g_last_error", "b2g_env_create", "b2g_env_load_yaml", "b2g_env_destroy", "b2g_env_add_object", "b2g_env_add_link",
    "b2g_env_add_polygon", "b2g_env_add_ball", "b2g_env_add_joint", "b2g_env_add_state", "b2g_model_create", "b2g_model_destroy",
    "b2g_model_get_info", "b2g_model_get_warnings", "b2g_model_get_object", "b2g_model_get_dof_limits", "b2g_model_get_bodies",
    "b2g_model_get_fixtures", "b2g_model_get_joints", "b2g_model_get_balls", "b2g_model_set_link_property",
    "b2g_model_get_link_property", "b2g_model_get_link", "b2g_model_find_link", "b2g_model_get_link_aabb", "b2g_model_get_object_aabb", "b2g_model_get_joint", "b2g_model_find_joint",
    "b2g_batch_get_ball_approximation",
    "b2g_batch_create", "b2g_batch_clone", "b2g_batch_destroy", "b2g_batch_set_params",
    "b2g_batch_status", "b2g_batch_stream", "b2g_batch_sync", "b2g_batch_set_state", "b2g_batch_set_state_dev",
    "b2g_batch_get_state", "b2g_batch_get_state_dev", "b2g_batch_set_object_pose", "b2g_batch_set_object_dofs", "b2g_batch_get_object_dofs", "b2g_batch_select_worlds", "b2g_batch_get_object_pose",
    "b2g_batch_set_to_rest", "b2g_batch_at_rest", "b2g_batch_set_link_enabled", "b2g_batch_set_link_property",
    "b2g_batch_get_link_property", "b2g_batch_objects_in_aabb",
    "b2g_batch_is_physically_feasible",
    "b2g_batch_set_target_velocity", "b2g_batch_set_target_velocity_dev", "b2g_batch_set_control_efforts",
    "b2g_batch_set_target_position", "b2g_batch_set_target_position_dev",
    "b2g_batch_set_control_efforts_dev", "b2g_batch_step", "b2g_batch_step_record",
    "b2g_batch_step_guarded", "b2g_batch_rollout_dev", "b2g_batch_rollout",
    "b2g_batch_get_link_states", "b2g_batch_get_bodies", "b2g_batch_get_fat_aabbs", "b2g_batch_get_joints", "b2g_batch_get_contacts",
    "b2g_batch_check_collision", "b2g_batch_check_collision_dev", "b2g_batch_save_state", "b2g_batch_restore_state", "b2g_batch_drop_state",
    "b2g_host_alloc", "b2g_host_free", "b2g_launch_count", "b2g_batch_take_kernel_ms",
    "b2g_multi_create", "b2g_multi_destroy", "b2g_multi_device_count", "b2g_multi_get_shard", "b2g_multi_rollout",
    "b2g_multi_gathered_dev", "b2g_multi_last_timing",
]


def lib():
    """Load libb2g.so (building it in-tree with nvcc if missing or stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libb2g.so")
    if os.environ.get("B2G_LIB"):  # tuning experiments only: a prebuilt variant of the library
        so = os.environ["B2G_LIB"]
    else:
        try:
            from . import build as _build
            so = _build.build()
        except Exception:
            if not os.path.exists(so):
                raise
    L = C.CDLL(so)
    vp, ip, fp = C.c_void_p, C.c_void_p, C.c_void_p
    L.b2g_last_error.restype = C.c_char_p
    L.b2g_env_create.argtypes = [C.c_float, fp, C.POINTER(vp)]
    L.b2g_env_load_yaml.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b2g_env_destroy.argtypes = [vp]
    L.b2g_env_destroy.restype = None
    L.b2g_env_add_object.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, fp, C.c_int, C.POINTER(C.c_int)]
    L.b2g_env_add_link.argtypes = [vp, C.c_int, C.c_int, C.c_char_p] + [C.c_float] * 5 + [C.POINTER(C.c_int)]
    L.b2g_env_add_polygon.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, C.c_int]
    L.b2g_env_add_ball.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, C.c_float]
    L.b2g_model_get_balls.argtypes = [vp, C.c_int, ip, fp]
    L.b2g_model_set_link_property.argtypes = [vp, C.c_int, C.c_int, C.c_float]
    L.b2g_model_get_link_property.argtypes = [vp, C.c_int, C.c_int, fp]
    L.b2g_batch_get_ball_approximation.argtypes = [vp, C.c_int, fp]
    L.b2g_env_add_joint.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, fp, C.c_int, C.c_char_p]
    L.b2g_env_add_state.argtypes = [vp, C.c_char_p, fp, fp, C.c_int]
    L.b2g_model_create.argtypes = [vp, C.POINTER(vp)]
    L.b2g_model_destroy.argtypes = [vp]
    L.b2g_model_destroy.restype = None
    L.b2g_model_get_info.argtypes = [vp, C.POINTER(_ModelInfo)]
    L.b2g_model_get_object.argtypes = [vp, C.c_int, C.POINTER(_ObjectInfo)]
    L.b2g_model_get_link.argtypes = [vp, C.c_int, C.POINTER(_LinkInfo)]
    L.b2g_model_find_link.argtypes = [vp, C.c_int, C.c_char_p, ip]
    L.b2g_model_get_link_aabb.argtypes = [vp, C.c_int, fp]
    L.b2g_model_get_object_aabb.argtypes = [vp, C.c_int, fp]
    L.b2g_model_get_joint.argtypes = [vp, C.c_int, C.POINTER(_JointInfo)]
    L.b2g_model_find_joint.argtypes = [vp, C.c_int, C.c_char_p, ip]
    L.b2g_model_get_warnings.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
    L.b2g_model_get_dof_limits.argtypes = [vp, fp]
    L.b2g_model_get_bodies.argtypes = [vp, fp]
    L.b2g_model_get_fixtures.argtypes = [vp, ip, fp]
    L.b2g_model_get_joints.argtypes = [vp, ip, fp]
    L.b2g_batch_create.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.POINTER(vp)]
    L.b2g_batch_clone.argtypes = [vp, C.POINTER(vp)]
    L.b2g_batch_destroy.argtypes = [vp]
    L.b2g_batch_destroy.restype = None
    L.b2g_batch_set_params.argtypes = [vp, C.c_float, C.c_int, C.c_int]
    L.b2g_batch_status.argtypes = [vp, ip]
    L.b2g_batch_stream.argtypes = [vp]
    L.b2g_batch_stream.restype = C.c_void_p
    L.b2g_batch_sync.argtypes = [vp]
    for n in ("b2g_batch_set_state", "b2g_batch_set_state_dev"):
        getattr(L, n).argtypes = [vp, fp, fp, C.c_int]
    for n in ("b2g_batch_get_state", "b2g_batch_get_state_dev"):
        getattr(L, n).argtypes = [vp, fp, fp]
    L.b2g_batch_set_object_dofs.argtypes = [vp, C.c_int, C.c_int, ip, C.c_int, fp]
    L.b2g_batch_get_object_dofs.argtypes = [vp, C.c_int, C.c_int, ip, C.c_int, fp]
    L.b2g_batch_select_worlds.argtypes = [vp, C.c_int64, C.c_int64]
    L.b2g_batch_set_object_pose.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_get_object_pose.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_set_to_rest.argtypes = [vp]
    L.b2g_batch_at_rest.argtypes = [vp, C.c_float, ip]
    L.b2g_batch_set_link_enabled.argtypes = [vp, C.c_int, vp, C.c_int]
    L.b2g_batch_set_link_property.argtypes = [vp, C.c_int, C.c_int, C.c_float]
    L.b2g_batch_get_link_property.argtypes = [vp, C.c_int, C.c_int, fp]
    L.b2g_batch_objects_in_aabb.argtypes = [vp, fp, C.c_int, vp]
    L.b2g_batch_is_physically_feasible.argtypes = [vp, ip]
    L.b2g_batch_set_target_velocity.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_set_target_velocity_dev.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_set_control_efforts.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_set_control_efforts_dev.argtypes = [vp, C.c_int, fp]
    L.b2g_batch_set_target_position.argtypes = [vp, C.c_int, fp, C.c_float]
    L.b2g_batch_set_target_position_dev.argtypes = [vp, C.c_int, fp, C.c_float]
    L.b2g_batch_step.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.b2g_batch_rollout_dev.argtypes = [vp, C.c_int, fp, C.c_int, C.c_int]
    L.b2g_batch_rollout.argtypes = [vp, C.c_int, fp, fp, C.c_int, fp, C.c_int, C.c_int, fp, fp]
    L.b2g_batch_step_record.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, fp, C.c_int]
    L.b2g_batch_step_guarded.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, ip, ip]
    L.b2g_batch_get_bodies.argtypes = [vp, fp]
    L.b2g_batch_get_link_states.argtypes = [vp, fp]
    L.b2g_batch_get_fat_aabbs.argtypes = [vp, fp]
    L.b2g_batch_get_joints.argtypes = [vp, ip, fp]
    L.b2g_batch_get_contacts.argtypes = [vp, ip, ip, fp, C.c_int]
    L.b2g_batch_check_collision.argtypes = [vp, ip, ip, fp, C.c_int]
    L.b2g_batch_check_collision_dev.argtypes = [vp, ip, ip, fp, C.c_int]
    for n in ("b2g_batch_save_state", "b2g_batch_restore_state", "b2g_batch_drop_state"):
        getattr(L, n).argtypes = [vp]
    L.b2g_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.b2g_host_free.argtypes = [vp]
    L.b2g_host_free.restype = None
    L.b2g_launch_count.restype = C.c_uint64
    L.b2g_batch_take_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.b2g_multi_create.argtypes = [vp, C.c_int64, ip, C.c_int, C.c_int, C.POINTER(vp)]
    L.b2g_multi_destroy.argtypes = [vp]
    L.b2g_multi_destroy.restype = None
    L.b2g_multi_device_count.argtypes = [vp]
    L.b2g_multi_get_shard.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.b2g_multi_rollout.argtypes = [vp, C.c_int, fp, fp, C.c_int, fp, C.c_int, C.c_int, fp, fp]
    L.b2g_multi_gathered_dev.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]
    L.b2g_multi_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _LIB = L
    return L


def _check(code):
    if code != 0:
        raise B2GError(code, lib().b2g_last_error().decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pinned_empty(shape, dtype=np.float32):
    """numpy array backed by page-locked host memory (cudaHostAlloc through the C ABI)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _check(lib().b2g_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    # the ctypes buffer is the base object of every view of the array: the page-locked block is released when the last
    # view is gone (the finalizer holds the address, not the array)
    weakref.finalize(buf, lib().b2g_host_free, C.c_void_p(p.value))
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Environment:
    """Box2DEnvironmentDescription (include/sim_env/Box2DIOUtils.h:64-70)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_yaml(cls, path):
        h = C.c_void_p()
        _check(lib().b2g_env_load_yaml(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def from_dict(cls, env):
        """Programmatic construction from the plain-dict description used by envgen / tests."""
        L = lib()
        h = C.c_void_p()
        b = _f32(env.get("world_bounds", [0, 0, 0, 0]))
        _check(L.b2g_env_create(float(env["scale"]), _ptr(b), C.byref(h)))
        self = cls(h)
        for is_robot, key in ((1, "robots"), (0, "objects")):
            for obj in env[key]:
                ba = obj.get("base_actuation")
                lim = None
                use_com = 1
                if ba:
                    lim = _f32(list(ba["translational_velocity_limits"]) + list(ba["translational_acceleration_limits"]) +
                               list(ba["rotational_velocity_limits"]) + list(ba["rotational_acceleration_limits"]))
                    use_com = int(ba.get("use_center_of_mass", True))
                oi = C.c_int()
                _check(L.b2g_env_add_object(h, obj["name"].encode(), is_robot, int(obj["static"]), _ptr(lim), use_com, C.byref(oi)))
                for link in obj["links"]:
                    li = C.c_int()
                    _check(L.b2g_env_add_link(h, is_robot, oi.value, link["name"].encode(), link["mass"], link["ground_friction"],
                                              link["ground_torque_integral"], link["contact_friction"], link["restitution"],
                                              C.byref(li)))
                    for poly in link["geometry"]:
                        p = _f32(poly)
                        _check(L.b2g_env_add_polygon(h, is_robot, oi.value, li.value, _ptr(p), len(p)))
                    for ball in link.get("balls") or []:
                        c = _f32(ball[:3])
                        _check(L.b2g_env_add_ball(h, is_robot, oi.value, li.value, _ptr(c), float(ball[3])))
                for j in obj["joints"]:
                    params = _f32(list(j["axis"]) + list(j["position_limits"]) + list(j["velocity_limits"]) +
                                  list(j["acceleration_limits"]) + [j["axis_orientation"]])
                    _check(L.b2g_env_add_joint(h, is_robot, oi.value, j["name"].encode(), j["link_a"].encode(),
                                               j["link_b"].encode(), _ptr(params), int(j["actuated"]), j["joint_type"].encode()))
        for name, st in env.get("states", {}).items():
            cfg = _f32(st["configuration"])
            vel = _f32(st["velocity"])
            _check(L.b2g_env_add_state(h, name.encode(), _ptr(cfg), _ptr(vel), len(cfg)))
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2g_env_destroy(self.h)
            self.h = None


# Box2DLink's dynamics parameters (Box2DWorld.cpp:471-566), numbered as b2g_link_property in include/b2g.h
LINK_PROPERTIES = dict(mass=0, ground_friction=1, ground_torque_integral=2, contact_friction=3, inertia=4,
                       com_inertia=5, ground_friction_limit_force=6, ground_friction_limit_torque=7)


class Model:
    """Immutable flattened model shared by the worlds of a batch (host only; no GPU needed)."""

    def __init__(self, env):
        if isinstance(env, (str, os.PathLike)):
            env = Environment.from_yaml(env)
        elif isinstance(env, dict):
            env = Environment.from_dict(env)
        self.env = env
        self.h = C.c_void_p()
        _check(lib().b2g_model_create(env.h, C.byref(self.h)))
        info = _ModelInfo()
        _check(lib().b2g_model_get_info(self.h, C.byref(info)))
        self.n_bodies, self.n_fixtures, self.n_joints = info.n_bodies, info.n_fixtures, info.n_joints
        self.n_objects, self.n_robots, self.nq, self.scale = info.n_objects, info.n_robots, info.nq, info.scale
        self.objects = []
        for i in range(self.n_objects):
            oi = _ObjectInfo()
            _check(lib().b2g_model_get_object(self.h, i, C.byref(oi)))
            self.objects.append(dict(name=oi.name.decode(), is_robot=bool(oi.is_robot), is_static=bool(oi.is_static),
                                     n_dofs=oi.n_dofs, dof_offset=oi.dof_offset, first_body=oi.first_body,
                                     n_links=oi.n_links, n_joints=oi.n_joints, n_balls=oi.n_balls))

    def warnings(self):
        """Loader diagnostics (what the reference sends to its logger while loading + accepted stale key names)."""
        buf = C.create_string_buffer(16384)
        n = C.c_int32()
        _check(lib().b2g_model_get_warnings(self.h, buf, len(buf), C.byref(n)))
        return [l for l in buf.value.decode().split("\n") if l]

    def object_index(self, name):
        for i, o in enumerate(self.objects):
            if o["name"] == name:
                return i
        raise KeyError(name)

    def dof_limits(self):
        out = np.zeros((self.nq, 6), np.float32)
        _check(lib().b2g_model_get_dof_limits(self.h, _ptr(out)))
        return out

    def body_model(self):
        out = np.zeros((self.n_bodies, 8), np.float32)
        _check(lib().b2g_model_get_bodies(self.h, _ptr(out)))
        return out

    def fixture_model(self):
        io = np.zeros((self.n_fixtures, 3), np.int32)
        fo = np.zeros((self.n_fixtures, 38), np.float32)
        _check(lib().b2g_model_get_fixtures(self.h, _ptr(io), _ptr(fo)))
        return io, fo

    def local_ball_approximation(self, obj):
        """Box2DLink::getLocalBallApproximation (:727-731) of every link of the object, links in name order:
        (bodies [n_balls], xyzr [n_balls, 4])."""
        obj = self.object_index(obj) if isinstance(obj, str) else obj
        n = self.objects[obj]["n_balls"]
        bodies = np.zeros(n, np.int32)
        xyzr = np.zeros((n, 4), np.float32)
        _check(lib().b2g_model_get_balls(self.h, obj, _ptr(bodies), _ptr(xyzr)))
        return bodies, xyzr

    def joint_model(self):
        io = np.zeros((self.n_joints, 6), np.int32)
        fo = np.zeros((self.n_joints, 15), np.float32)
        _check(lib().b2g_model_get_joints(self.h, _ptr(io), _ptr(fo)))
        return io, fo

    def link(self, body):
        """Name and place of the link with body index `body` (what contacts report as link_a / link_b)."""
        li = _LinkInfo()
        _check(lib().b2g_model_get_link(self.h, int(body), C.byref(li)))
        return dict(name=li.name.decode(), object=li.object, is_base_link=bool(li.is_base_link), is_static=bool(li.is_static),
                    parent_joint=li.parent_joint, n_fixtures=li.n_fixtures)

    def find_link(self, obj, name):
        """Box2DObject::getLink(name): body index of the link."""
        obj = self.object_index(obj) if isinstance(obj, str) else obj
        out = np.zeros(1, np.int32)
        _check(lib().b2g_model_find_link(self.h, obj, name.encode(), _ptr(out)))
        return int(out[0])

    def link_aabb(self, body):
        """Box2DLink::getLocalBoundingBox: min x, min y, max x, max y of the link's polygons (user units)."""
        out = np.zeros(4, np.float32)
        _check(lib().b2g_model_get_link_aabb(self.h, int(body), _ptr(out)))
        return out

    def object_aabb(self, obj):
        """Box2DObject::getLocalAABB."""
        obj = self.object_index(obj) if isinstance(obj, str) else obj
        out = np.zeros(4, np.float32)
        _check(lib().b2g_model_get_object_aabb(self.h, obj, _ptr(out)))
        return out

    def joint(self, joint):
        ji = _JointInfo()
        _check(lib().b2g_model_get_joint(self.h, int(joint), C.byref(ji)))
        return dict(name=ji.name.decode(), kind=ji.kind, object=ji.object, body_a=ji.body_a, body_b=ji.body_b,
                    dof_index=ji.dof_index, actuated=bool(ji.actuated), limited=bool(ji.limited))

    def find_joint(self, obj, name):
        """Box2DObject::getJoint(name): model joint index of the revolute joint."""
        obj = self.object_index(obj) if isinstance(obj, str) else obj
        out = np.zeros(1, np.int32)
        _check(lib().b2g_model_find_joint(self.h, obj, name.encode(), _ptr(out)))
        return int(out[0])

    def set_link_property(self, body, prop, value):
        """Box2DLink's setters (:478-547) applied to the model: batches created afterwards start with the new value."""
        _check(lib().b2g_model_set_link_property(self.h, int(body), LINK_PROPERTIES[prop], float(value)))

    def get_link_property(self, body, prop):
        out = np.zeros(1, np.float32)
        _check(lib().b2g_model_get_link_property(self.h, int(body), LINK_PROPERTIES[prop], _ptr(out)))
        return float(out[0])

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2g_model_destroy(self.h)
            self.h = None


class BatchBox2DWorld:
    """N independent Box2DWorld instances stepped together on one B200.

    Method names and argument meaning follow sim_env::Box2DWorld; every array gains a leading world axis.
    """

    def __init__(self, n_worlds=1, device=0, max_contacts=0):
        self.n_worlds = int(n_worlds)
        self.device = device
        self.max_contacts = max_contacts
        self.model = None
        self.h = None

    # -- loadWorld(path) / loadWorld(env_desc)  (Box2DWorld.cpp:3040-3054)
    def loadWorld(self, env):
        self._destroy()
        self.model = env if isinstance(env, Model) else Model(env)
        self.h = C.c_void_p()
        _check(lib().b2g_batch_create(self.model.h, self.n_worlds, self.device, self.max_contacts, C.byref(self.h)))
        self.nq = self.model.nq
        return self

    def clone(self):
        """Box2DWorld::clone (:3021-3038): same environment and step parameters, setWorldState(getWorldState()) incl. the
        poses of static objects, fresh caches."""
        other = BatchBox2DWorld(self.n_worlds, self.device, self.max_contacts)
        other.model = self.model
        other.nq = self.nq
        other._dt, other._vs, other._ps = self.getPhysicsTimeStep(), self.getVelocitySteps(), self.getPositionSteps()
        other.h = C.c_void_p()
        _check(lib().b2g_batch_clone(self.h, C.byref(other.h)))
        return other

    def _destroy(self):
        if getattr(self, "h", None):
            lib().b2g_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        self._destroy()

    # -- physics parameters (:3340-3368)
    def setPhysicsTimeStep(self, dt):
        self._dt = float(dt)
        self._push_params()

    def setVelocitySteps(self, n):
        self._vs = int(n)
        self._push_params()

    def setPositionSteps(self, n):
        self._ps = int(n)
        self._push_params()

    def getPhysicsTimeStep(self):
        return getattr(self, "_dt", 0.01)

    def getVelocitySteps(self):
        return getattr(self, "_vs", 15)

    def getPositionSteps(self):
        return getattr(self, "_ps", 15)

    def _push_params(self):
        _check(lib().b2g_batch_set_params(self.h, self.getPhysicsTimeStep(), self.getVelocitySteps(), self.getPositionSteps()))

    # -- stepping (:3256-3283)
    def stepPhysics(self, steps=1, execute_controllers=True, allow_sleeping=True):
        _check(lib().b2g_batch_step(self.h, int(steps), int(execute_controllers), int(allow_sleeping)))

    def stepPhysicsRecord(self, steps=1, execute_controllers=True, allow_sleeping=True, max_per_world=32):
        """stepPhysics(std::vector<Contact>& contacts, steps, ...) (:3290-3310): (counts, ints, floats) per world."""
        counts = np.zeros(self.n_worlds, np.int32)
        io = np.zeros((self.n_worlds, max_per_world, 4), np.int32)
        fo = np.zeros((self.n_worlds, max_per_world, 6), np.float32)
        _check(lib().b2g_batch_step_record(self.h, int(steps), int(execute_controllers), int(allow_sleeping), _ptr(counts),
                                           _ptr(io), _ptr(fo), max_per_world))
        return counts, io, fo

    def stepPhysicsGuarded(self, allowed_pairs, steps=1, execute_controllers=True, allow_sleeping=True):
        """stepPhysics(callback, steps, ...) (:3312-3333) with the callback as an [n_bodies, n_bodies] table (0 = abort).
        Returns (completed [n_worlds] bool, steps_done [n_worlds])."""
        nb = self.model.n_bodies
        table = np.ascontiguousarray(allowed_pairs, dtype=np.uint8)
        if table.shape != (nb, nb):
            raise B2GError(1, "allowed_pairs must be [%d, %d]" % (nb, nb))
        completed = np.zeros(self.n_worlds, np.int32)
        done = np.zeros(self.n_worlds, np.int32)
        _check(lib().b2g_batch_step_guarded(self.h, int(steps), int(execute_controllers), int(allow_sleeping), _ptr(table),
                                            _ptr(completed), _ptr(done)))
        return completed.astype(bool), done

    # -- state (:3462-3541)
    def getWorldState(self):
        q = np.zeros((self._rows(), self.nq), np.float32)
        qd = np.zeros((self._rows(), self.nq), np.float32)
        _check(lib().b2g_batch_get_state(self.h, _ptr(q), _ptr(qd)))
        return q, qd

    def setWorldState(self, q, qd, reset_caches=False):
        q = _f32(q)
        qd = _f32(qd)
        if q.shape != (self._rows(), self.nq) or qd.shape != (self._rows(), self.nq):
            raise B2GError(1, "[sim_env::BatchBox2DWorld::setWorldState] state arrays must be [n_worlds, %d]" % self.nq)
        _check(lib().b2g_batch_set_state(self.h, _ptr(q), _ptr(qd), int(reset_caches)))

    def saveState(self):
        _check(lib().b2g_batch_save_state(self.h))

    def restoreState(self):
        code = lib().b2g_batch_restore_state(self.h)
        if code == 1 and b"stack is empty" in lib().b2g_last_error():
            return False
        _check(code)
        return True

    def dropState(self):
        _check(lib().b2g_batch_drop_state(self.h))

    def setObjectPose(self, obj, pose):
        obj = self.model.object_index(obj) if isinstance(obj, str) else obj
        pose = _f32(np.broadcast_to(_f32(pose), (self._rows(), 3)))
        _check(lib().b2g_batch_set_object_pose(self.h, obj, _ptr(pose)))

    def getObjectPose(self, obj):
        obj = self.model.object_index(obj) if isinstance(obj, str) else obj
        pose = np.zeros((self._rows(), 3), np.float32)
        _check(lib().b2g_batch_get_object_pose(self.h, obj, _ptr(pose)))
        return pose

    # ---- per-object DOF access (Box2DObject::get/setDOFPositions, get/setDOFVelocities with an index subset)
    def _dofs(self, obj, indices):
        obj = self.model.object_index(obj) if isinstance(obj, str) else obj
        ix = None if indices is None else np.ascontiguousarray(indices, np.int32)
        return obj, ix, (self.model.objects[obj]["n_dofs"] if ix is None else len(ix))

    def getDOFPositions(self, obj, indices=None, velocities=False):
        obj, ix, n = self._dofs(obj, indices)
        out = np.zeros((self._rows(), n), np.float32)
        _check(lib().b2g_batch_get_object_dofs(self.h, obj, int(velocities), _ptr(ix), 0 if ix is None else n, _ptr(out)))
        return out

    def getDOFVelocities(self, obj, indices=None):
        return self.getDOFPositions(obj, indices, velocities=True)

    def setDOFPositions(self, obj, values, indices=None, velocities=False):
        obj, ix, n = self._dofs(obj, indices)
        v = _f32(np.broadcast_to(_f32(values), (self._rows(), n)))
        _check(lib().b2g_batch_set_object_dofs(self.h, obj, int(velocities), _ptr(ix), 0 if ix is None else n, _ptr(v)))

    def setDOFVelocities(self, obj, values, indices=None):
        self.setDOFPositions(obj, values, indices, velocities=True)

    def selectWorlds(self, first=0, count=0):
        """Restrict the per-world calls to worlds [first, first + count) (count <= 0: the whole batch again)."""
        _check(lib().b2g_batch_select_worlds(self.h, int(first), int(count)))
        self._sel = (int(first), int(count)) if count > 0 else None

    def _rows(self):
        sel = getattr(self, "_sel", None)
        return sel[1] if sel else self.n_worlds

    def getBallApproximation(self, obj):
        """Box2DObject::getBallApproximation (:1754-1763) in every world: [n_worlds, n_balls, 4] = centre, radius."""
        obj = self.model.object_index(obj) if isinstance(obj, str) else obj
        out = np.zeros((self._rows(), self.model.objects[obj]["n_balls"], 4), np.float32)
        _check(lib().b2g_batch_get_ball_approximation(self.h, obj, _ptr(out)))
        return out

    def getObjects(self, aabb, exclude_robots=True):
        """Box2DWorld::getObjects(aabb, ...) (:3207-3254): bool [n_worlds, n_objects] (model object order)."""
        box = _f32(aabb)
        hits = np.zeros((self._rows(), self.model.n_objects), np.uint8)
        _check(lib().b2g_batch_objects_in_aabb(self.h, _ptr(box), int(exclude_robots), _ptr(hits)))
        return hits.astype(bool)

    def isPhysicallyFeasible(self):
        """Box2DWorld::isPhysicallyFeasible (:3370-3400), one bool per world."""
        out = np.zeros(self._rows(), np.int32)
        _check(lib().b2g_batch_is_physically_feasible(self.h, _ptr(out)))
        return out.astype(bool)

    def setToRest(self):
        """Box2DWorld::setToRest (:3550-3558)."""
        _check(lib().b2g_batch_set_to_rest(self.h))

    def atRest(self, threshold=0.0001):
        """Box2DWorld::atRest (:3435-3452), one bool per world."""
        out = np.zeros(self._rows(), np.int32)
        _check(lib().b2g_batch_at_rest(self.h, float(threshold), _ptr(out)))
        return out.astype(bool)

    def setLinkEnabled(self, body, enabled):
        """Box2DLink::setEnabled (:802-822) for the link with body index `body`; enabled: bool or [n_worlds] array."""
        if np.isscalar(enabled):
            _check(lib().b2g_batch_set_link_enabled(self.h, int(body), None, int(bool(enabled))))
        else:
            e = np.ascontiguousarray(enabled, dtype=np.uint8)
            if e.shape != (self._rows(),):
                raise B2GError(1, "enabled must be a bool or an [n_worlds] array")
            _check(lib().b2g_batch_set_link_enabled(self.h, int(body), _ptr(e), 0))

    LINK_PROPERTIES = LINK_PROPERTIES

    def setLinkProperty(self, body, prop, value):
        """Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction
        (:478-547) for the link with body index `body` in every SELECTED world.  The whole batch, or -- for per-world
        parameters (domain randomisation) -- a selection that covers whole 32-world tiles (selectWorlds(first, count) with
        first % 32 == 0 and count % 32 == 0 or first + count == n_worlds): the worlds of a tile share one parameter set."""
        _check(lib().b2g_batch_set_link_property(self.h, int(body), self.LINK_PROPERTIES[prop], float(value)))

    def getLinkProperty(self, body, prop):
        out = np.zeros(1, np.float32)
        _check(lib().b2g_batch_get_link_property(self.h, int(body), self.LINK_PROPERTIES[prop], _ptr(out)))
        return float(out[0])

    # -- controller (Box2DController.h:23)
    def setTargetVelocity(self, robot, velocity):
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        nd = self.model.objects[robot]["n_dofs"]
        v = _f32(np.broadcast_to(_f32(velocity), (self._rows(), nd)))
        if v.shape[-1] != nd:
            raise B2GError(1, "[sim_env::Box2DRobotVelocityController::setTargetVelocity]"
                              "Provided target velocity has invalid dimension.")
        _check(lib().b2g_batch_set_target_velocity(self.h, robot, _ptr(v)))

    def setTargetPosition(self, robot, position, kp=5.0):
        """Position controller on top of the velocity controller (b2g_batch_set_target_position: v_target = kp * (q_target - q)
        every control step, heading error wrapped; the reference's sim_env::RobotPositionController is not in the tree)."""
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        nd = self.model.objects[robot]["n_dofs"]
        q = _f32(np.broadcast_to(_f32(position), (self._rows(), nd)))
        _check(lib().b2g_batch_set_target_position(self.h, robot, _ptr(q), C.c_float(kp)))

    def setControlEfforts(self, robot, efforts):
        """Box2DRobot::setController (:2312-2317) with a callback that returns `efforts` ([n_worlds, n_dofs]) on every step:
        they go to commandEfforts (:2319-2379).  The hook for user control laws."""
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        nd = self.model.objects[robot]["n_dofs"]
        e = _f32(np.broadcast_to(_f32(efforts), (self._rows(), nd)))
        _check(lib().b2g_batch_set_control_efforts(self.h, robot, _ptr(e)))

    # -- collision (:3846-3859 through Box2DCollisionChecker::updateContactCache :2880-2911)
    def checkCollision(self, max_per_world=32):
        counts = np.zeros(self._rows(), np.int32)
        io = np.zeros((self._rows(), max_per_world, 4), np.int32)
        fo = np.zeros((self._rows(), max_per_world, 6), np.float32)
        _check(lib().b2g_batch_check_collision(self.h, _ptr(counts), _ptr(io), _ptr(fo), max_per_world))
        return counts, io, fo

    # -- parity probes
    def getLinkStates(self):
        """Box2DLink::getPose / getVelocityVector / getCenterOfMass of every link: [n_worlds, n_bodies, 8] =
        pose x y theta, velocity x y omega, centre of mass x y (user units; row 0 is the ground)."""
        out = np.zeros((self._rows(), self.model.n_bodies, 8), np.float32)
        _check(lib().b2g_batch_get_link_states(self.h, _ptr(out)))
        return out

    def bodies(self):
        out = np.zeros((self._rows(), self.model.n_bodies, 16), np.float32)
        _check(lib().b2g_batch_get_bodies(self.h, _ptr(out)))
        return out

    def fat_aabbs(self):
        out = np.zeros((self._rows(), self.model.n_fixtures, 4), np.float32)
        _check(lib().b2g_batch_get_fat_aabbs(self.h, _ptr(out)))
        return out

    def joints(self):
        io = np.zeros((self._rows(), self.model.n_joints, 6), np.int32)
        fo = np.zeros((self._rows(), self.model.n_joints, 15), np.float32)
        _check(lib().b2g_batch_get_joints(self.h, _ptr(io), _ptr(fo)))
        return io, fo

    def contacts(self, max_per_world=32):
        counts = np.zeros(self._rows(), np.int32)
        io = np.zeros((self._rows(), max_per_world, 8), np.int32)
        fo = np.zeros((self._rows(), max_per_world, 12), np.float32)
        _check(lib().b2g_batch_get_contacts(self.h, _ptr(counts), _ptr(io), _ptr(fo), max_per_world))
        return counts, io, fo

    def status(self):
        out = np.zeros(self._rows(), np.int32)
        _check(lib().b2g_batch_status(self.h, _ptr(out)))
        return out

    def take_kernel_ms(self):
        ms = C.c_float()
        n = C.c_int32()
        _check(lib().b2g_batch_take_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- device-pointer entry points (torch / CUDA interop; pointers are raw ints)
    def set_state_dev(self, q_ptr, qd_ptr, reset_caches=True):
        _check(lib().b2g_batch_set_state_dev(self.h, C.c_void_p(q_ptr), C.c_void_p(qd_ptr), int(reset_caches)))

    def get_state_dev(self, q_ptr, qd_ptr):
        _check(lib().b2g_batch_get_state_dev(self.h, C.c_void_p(q_ptr), C.c_void_p(qd_ptr)))

    def rollout(self, robot, targets, steps_per_target, q=None, qd=None, reset_caches=True, q_out=None, qd_out=None):
        """Planner "propagate" as one host-pointer call: optional setWorldState(q, qd), then for every targets[t]
        ([n_targets, n_worlds, n_dofs]) setTargetVelocity + stepPhysics(steps_per_target), then getWorldState."""
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        tg = _f32(targets)
        nd = self.model.objects[robot]["n_dofs"]
        if tg.ndim != 3 or tg.shape[1:] != (self.n_worlds, nd):
            raise B2GError(1, "targets must be [n_targets][n_worlds][n_dofs]")
        if q is not None:
            q, qd = _f32(q), _f32(qd)
            if q.shape != (self.n_worlds, self.nq) or qd.shape != q.shape:
                raise B2GError(1, "state arrays must be [n_worlds][nq]")
        q_out = np.empty((self.n_worlds, self.nq), np.float32) if q_out is None else q_out
        qd_out = np.empty((self.n_worlds, self.nq), np.float32) if qd_out is None else qd_out
        _check(lib().b2g_batch_rollout(self.h, robot, _ptr(q), _ptr(qd), int(reset_caches), _ptr(tg), tg.shape[0],
                                       int(steps_per_target), _ptr(q_out), _ptr(qd_out)))
        return q_out, qd_out

    def rollout_host(self, robot, q_ptr, qd_ptr, targets_ptr, n_targets, steps_per_target, q_out_ptr, qd_out_ptr, reset_caches=True):
        """b2g_batch_rollout on raw host addresses (pinned buffers of the caller)."""
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        _check(lib().b2g_batch_rollout(self.h, robot, C.c_void_p(q_ptr), C.c_void_p(qd_ptr), int(reset_caches), C.c_void_p(targets_ptr),
                                       int(n_targets), int(steps_per_target), C.c_void_p(q_out_ptr), C.c_void_p(qd_out_ptr)))

    def rollout_dev(self, robot, targets_ptr, n_targets, steps_per_target):
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        _check(lib().b2g_batch_rollout_dev(self.h, robot, C.c_void_p(targets_ptr), int(n_targets), int(steps_per_target)))

    def check_collision_dev(self, counts_ptr, io_ptr=None, fo_ptr=None, max_per_world=0):
        _check(lib().b2g_batch_check_collision_dev(self.h, C.c_void_p(counts_ptr), C.c_void_p(io_ptr) if io_ptr else None,
                                                   C.c_void_p(fo_ptr) if fo_ptr else None, int(max_per_world)))

    def stream(self):
        return lib().b2g_batch_stream(self.h)

    def sync(self):
        _check(lib().b2g_batch_sync(self.h))


class MultiGpuBatch:
    """b2g_multi_*: one batch per GPU of this node inside ONE process (one host thread per device), contiguous world slices,
    ncclAllGather of the final rollout states over NVLink.  Mirrors what a C++ planner gets from include/b2g.h."""

    def __init__(self, model, n_worlds, devices, max_contacts=0):
        self.model = model if isinstance(model, Model) else Model(model)
        self.n_worlds = int(n_worlds)
        self.devices = [int(d) for d in devices]
        dev = np.asarray(self.devices, np.int32)
        h = C.c_void_p()
        _check(lib().b2g_multi_create(self.model.h, self.n_worlds, _ptr(dev), len(self.devices), int(max_contacts), C.byref(h)))
        self.h = h
        self.nq = self.model.nq

    def shard(self, i):
        """(batch handle, first world, world count) of device i."""
        b, first, count = C.c_void_p(), C.c_int64(), C.c_int64()
        _check(lib().b2g_multi_get_shard(self.h, int(i), C.byref(b), C.byref(first), C.byref(count)))
        return b, first.value, count.value

    def setObjectPose(self, obj, pose):
        """b2g_batch_set_object_pose on every shard; pose: [n_worlds, 3] for the whole job."""
        obj = self.model.object_index(obj) if isinstance(obj, str) else obj
        pose = _f32(np.broadcast_to(_f32(pose), (self.n_worlds, 3)))
        for i in range(len(self.devices)):
            b, first, count = self.shard(i)
            _check(lib().b2g_batch_set_object_pose(b, obj, _ptr(np.ascontiguousarray(pose[first:first + count]))))

    def rollout(self, robot, targets, steps_per_target, q=None, qd=None, reset_caches=True, q_out=None, qd_out=None):
        robot = self.model.object_index(robot) if isinstance(robot, str) else robot
        tg = _f32(targets)
        if q is not None:
            q, qd = _f32(q), _f32(qd)
        q_out = np.empty((self.n_worlds, self.nq), np.float32) if q_out is None else q_out
        qd_out = np.empty((self.n_worlds, self.nq), np.float32) if qd_out is None else qd_out
        _check(lib().b2g_multi_rollout(self.h, robot, _ptr(q), _ptr(qd), int(reset_caches), _ptr(tg), tg.shape[0],
                                       int(steps_per_target), _ptr(q_out), _ptr(qd_out)))
        return q_out, qd_out

    def rollout_host(self, robot, q_ptr, qd_ptr, targets_ptr, n_targets, steps_per_target, q_out_ptr, qd_out_ptr, reset_caches=True):
        _check(lib().b2g_multi_rollout(self.h, robot, C.c_void_p(q_ptr), C.c_void_p(qd_ptr), int(reset_caches), C.c_void_p(targets_ptr),
                                       int(n_targets), int(steps_per_target), C.c_void_p(q_out_ptr), C.c_void_p(qd_out_ptr)))

    def last_timing(self):
        lo, hi, g = C.c_float(), C.c_float(), C.c_float()
        _check(lib().b2g_multi_last_timing(self.h, C.byref(lo), C.byref(hi), C.byref(g)))
        return {"kernel_ms_min": lo.value, "kernel_ms_max": hi.value, "gather_ms": g.value}

    def gathered_dev(self, i):
        """(q_all ptr, qd_all ptr, rows per device) of the gathered states on device i."""
        qa, qda, rows = C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(lib().b2g_multi_gathered_dev(self.h, int(i), C.byref(qa), C.byref(qda), C.byref(rows)))
        return qa.value, qda.value, rows.value

    def close(self):
        if getattr(self, "h", None):
            lib().b2g_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
