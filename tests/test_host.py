"""CPU-side tests of the product: YAML loader, model builder (against the oracle's independent builder), C ABI
surface, and loud failure without a GPU.  No compute entry point is exercised here."""
import ctypes
import os
import re

import numpy as np
import pytest

from parity_utils import ENV_2, ENV_3, ENV_A, ENV_B, env_variant, load_env_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(b2g):
    header = open(os.path.join(ROOT, "include", "b2g.h")).read()
    declared = set(re.findall(r"^(?:b2g_status|void\*?|const char\*|uint64_t|int32_t)\s+(b2g_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no declarations found"
    L = b2g.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libb2g.so does not export %s" % name
    from box2d_sim_env_b200.batch_world import ABI_SYMBOLS
    assert declared == set(ABI_SYMBOLS)


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B])
def test_model_matches_oracle_builder(b2g, oracle, env_name):
    m = b2g.Model(b2g.data_path(env_name))
    w = oracle.OracleWorld(load_env_dict(env_name))
    assert (m.n_bodies, m.n_fixtures, m.n_joints, m.nq) == (w.n_bodies, w.n_proxies, w.n_joints, w.nq)
    assert np.array_equal(m.body_model(), w.body_model())          # bit-exact mass data
    mi, mf = m.fixture_model()
    oi, of = w.fixture_model()
    assert np.array_equal(mi, oi) and np.array_equal(mf, of)       # bit-exact hulls / normals / densities
    ji, jf = m.joint_model()
    oji, ojf = w.joints()
    assert np.array_equal(ji[:, [0, 1, 2, 4, 5]], oji[:, [0, 1, 2, 4, 5]])
    assert np.array_equal(jf[:, 6:], ojf[:, 6:])
    assert [o["name"] for o in m.objects] == [o["name"] for o in w.objects]
    assert [o["n_dofs"] for o in m.objects] == [o["n_dofs"] for o in w.objects]


@pytest.mark.parametrize("variant,ref_file", [(ENV_2, "test_env2.yaml"), (ENV_3, "test_env3.yaml")])
def test_reference_env2_env3(b2g, oracle, variant, ref_file):
    """The reference's other two environments (dynamic hinged obj2, frictionless fingers, no world_bounds key): the variants
    the parity tests use build the same model as the reference's own files read by the product's C++ loader (stale
    trans_friction / rot_friction keys, ignored max_torque key) and by the oracle's PyYAML loader."""
    env = env_variant(variant)
    m = b2g.Model(env)
    w = oracle.OracleWorld(env)
    assert np.array_equal(m.body_model(), w.body_model())
    assert np.array_equal(m.fixture_model()[1], w.fixture_model()[1])
    assert [(o["name"], o["n_dofs"], o["is_static"]) for o in m.objects][1:] == [("obj1", 3, False), ("obj2", 4, False)]
    ref = os.path.join("/root/reference/test_data", ref_file)
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this machine")
    r = b2g.Model(ref)                                    # product loader on the reference's own file
    assert (r.n_bodies, r.n_fixtures, r.n_joints, r.nq) == (m.n_bodies, m.n_fixtures, m.n_joints, m.nq)
    assert np.array_equal(r.body_model(), m.body_model())
    assert np.array_equal(r.fixture_model()[1], m.fixture_model()[1])
    assert np.array_equal(r.joint_model()[0], m.joint_model()[0]) and np.array_equal(r.joint_model()[1], m.joint_model()[1])
    assert np.array_equal(r.dof_limits(), m.dof_limits())
    ow = oracle.OracleWorld(oracle.load_env_yaml(ref))    # oracle loader on the reference's own file
    q0, qd0 = ow.get_state()
    q1, qd1 = w.get_state()
    assert np.array_equal(q0, q1) and np.array_equal(qd0, qd1)


def test_programmatic_env_equals_yaml(b2g):
    a = b2g.Model(b2g.data_path(ENV_A))
    b = b2g.Model(load_env_dict(ENV_A))       # dict -> b2g_env_* builder calls
    assert np.array_equal(a.body_model(), b.body_model())
    assert np.array_equal(a.fixture_model()[1], b.fixture_model()[1])
    assert np.array_equal(a.joint_model()[1], b.joint_model()[1])
    assert np.array_equal(a.dof_limits(), b.dof_limits())


def test_dof_layout(b2g):
    m = b2g.Model(b2g.data_path(ENV_A))
    assert [(o["name"], o["n_dofs"], o["dof_offset"]) for o in m.objects] == [("my_robot", 7, 0), ("obj1", 3, 7), ("obj2", 1, 10)]
    lim = m.dof_limits()
    assert lim[0, 2:].tolist() == pytest.approx([-2, 2, -2, 2])
    assert lim[2].tolist() == pytest.approx([-np.pi, np.pi, -1.54, 1.54, -1.4, 1.4])
    assert lim[4].tolist() == pytest.approx([-1.05, 1.05, -10, 10, -30, 30])
    assert lim[10].tolist() == pytest.approx([-1.57, 1.57, -10, 10, -1, 1])


def test_yaml_stale_keys_are_aliased(b2g, tmp_path):
    """The reference data says trans_friction / rot_friction, its decoder reads ground_friction /
    ground_torque_integral: both spellings must load to the same model."""
    src = b2g.data_path("")
    import shutil
    dst = tmp_path / "data"
    shutil.copytree(src, dst)
    for sub in ("objects", "robots"):
        for f in os.listdir(dst / sub):
            p = dst / sub / f
            t = open(p).read().replace("ground_friction", "trans_friction").replace("ground_torque_integral", "rot_friction")
            open(p, "w").write(t)
    a = b2g.Model(b2g.data_path(ENV_A))
    b = b2g.Model(str(dst / ENV_A))
    assert np.array_equal(a.joint_model()[1], b.joint_model()[1])


def test_loader_warnings_are_reported(b2g):
    """The reference's own test_data still uses trans_friction / rot_friction and test_env2.yaml has no world_bounds: the
    loader accepts them and says so (the reference would log through sim_env::Logger)."""
    assert b2g.Model(b2g.data_path(ENV_A)).warnings() == []
    ref = "/root/reference/test_data/test_env2.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this machine")
    w = b2g.Model(ref).warnings()
    assert any("world bounds" in x.lower() for x in w)
    assert any("trans_friction" in x for x in w) and any("rot_friction" in x for x in w)


def test_yaml_errors_are_loud(b2g, tmp_path):
    with pytest.raises(b2g.B2GError) as e:
        b2g.Model(str(tmp_path / "missing.yaml"))
    assert e.value.code == 2
    bad = tmp_path / "bad.yaml"
    bad.write_text("scale: 10.0\nrobots:\n  - name: r\n    filename: nothing.yaml\n")
    with pytest.raises(b2g.B2GError):
        b2g.Model(str(bad))


def test_prismatic_joints_follow_the_reference(b2g, oracle):
    """joint_type: prismatic builds a b2PrismaticJoint for an object (Box2DWorld.cpp:920-940: anchor not scaled, limits scaled,
    maxMotorForce = min |acceleration limits|), with the model records the oracle builds; what the reference cannot do is
    refused when the model is built: a robot with a prismatic joint (setControlTorque throws, :1204-1208) and a `states` entry
    for an object that has one (updateBodyVelocities -> getGlobalAxisPosition throws, :789, :1282-1285)."""
    from parity_utils import env_prismatic
    env = env_prismatic(actuated=True)
    m = b2g.Model(env)
    w = oracle.OracleWorld(oracle.OracleEnv(env))
    assert (m.n_bodies, m.n_joints, m.nq) == (w.n_bodies, w.n_joints, w.nq) == (8, 12, 11)
    ji, jf = m.joint_model()
    oji, ojf = w.joints()
    assert ji[-1][0] == 2 and np.array_equal(ji[:, [0, 1, 2, 4, 5]], oji[:, [0, 1, 2, 4, 5]])
    assert np.array_equal(jf[:, 8:], ojf[:, 8:])       # anchors (3.5, 0) unscaled, reference angle, translation limits -1 .. 2
    assert list(jf[-1][8:10]) == [3.5, 0.0] and list(jf[-1][13:15]) == [-1.0, 2.0]
    env = load_env_dict(ENV_A)
    env["robots"][0]["joints"][0]["joint_type"] = "prismatic"
    with pytest.raises(b2g.B2GError) as e:
        b2g.Model(env)
    assert e.value.code == 4 and "setControlTorque" in str(e.value)
    env = env_prismatic()
    env["states"]["drawer"] = dict(configuration=[1.0, 1.0, 0.0, 0.0], velocity=[0.0] * 4)
    with pytest.raises(b2g.B2GError) as e:
        b2g.Model(env)
    assert e.value.code == 4
    # an unknown joint type is reported and treated as revolute (:893-898)
    env = load_env_dict(ENV_A)
    env["robots"][0]["joints"][1]["joint_type"] = "helical"
    m = b2g.Model(env)
    assert any("Unknown joint type helical" in x for x in m.warnings())


def test_maximum_model_sizes(b2g):
    """Capacity limits of a world (64 bodies / 64 broad-phase proxies, model.h): the largest model builds, one more object
    fails loudly at model creation, never at step time."""
    from box2d_sim_env_b200 import envgen
    m = b2g.Model(envgen.clutter_env(n_objects=57))       # ground + 5 robot links + 57 = 63 bodies, 1 + 6 + 57 = 64 proxies
    assert (m.n_bodies, m.n_fixtures) == (63, 64)
    with pytest.raises(b2g.B2GError) as e:
        b2g.Model(envgen.clutter_env(n_objects=58))
    assert "too many" in str(e.value)


def test_kinematic_structure_is_validated(b2g):
    env = load_env_dict(ENV_A)
    env["robots"][0]["joints"][0]["link_b"] = "nonexistent"
    with pytest.raises(b2g.B2GError):
        b2g.Model(env)


def test_no_cpu_fallback(b2g):
    """Without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(b2g.B2GError) as e:
        b2g.BatchBox2DWorld(4).loadWorld(b2g.data_path(ENV_A))
    assert e.value.code == 3
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under box2d_sim_env_b200/ may reference it."""
    pkg = os.path.join(ROOT, "box2d_sim_env_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "libb2o" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_yaml_formatting_variants_match_pyyaml(b2g, oracle, tmp_path):
    """The product's own YAML-subset reader against PyYAML (the oracle's loader) on the formatting hand-written environment
    files use: comments, quoted names, CRLF line ends, flow sequences that run over several lines, odd indentation."""
    data = os.path.dirname(b2g.data_path(ENV_A))
    import shutil
    shutil.copytree(data, tmp_path / "data")
    obj = tmp_path / "data" / "objects" / "test_object.yaml"
    text = obj.read_text()
    first_poly = re.search(r"- \[[^\]]*\]", text).group(0)
    inner = first_poly[3:-1].split(",")
    multi = "- [" + ",".join(inner[:3]) + ",   # a comment inside the sequence\n          " + ",".join(inner[3:]) + "\n        ]"
    text = text.replace(first_poly, multi, 1)
    text = re.sub(r"name: (\S+)", r'name: "\1"', text, count=1)
    obj.write_text(text.replace("\n", "\r\n"))
    env_file = tmp_path / "data" / ENV_A
    env_file.write_text("# leading comment\n---\n" + env_file.read_text().replace("scale: 10.0", "scale:    10.0   # trailing comment"))
    path = str(env_file)
    m = b2g.Model(path)
    w = oracle.OracleWorld(oracle.load_env_yaml(path))
    ref = b2g.Model(b2g.data_path(ENV_A))
    assert np.array_equal(m.body_model(), w.body_model()) and np.array_equal(m.body_model(), ref.body_model())
    assert np.array_equal(m.fixture_model()[1], w.fixture_model()[1])
    assert [o["name"] for o in m.objects] == [o["name"] for o in ref.objects]
    bad = tmp_path / "data" / "objects" / "test_object2.yaml"
    bad.write_text(bad.read_text().replace("]", "", 1))           # a sequence that never closes
    with pytest.raises(b2g.B2GError):
        b2g.Model(path)


def test_c_header_is_plain_c(tmp_path):
    """include/b2g.h is the C ABI: it has to compile as C99 (no C++-isms, no CUDA / torch types)."""
    import subprocess
    src = tmp_path / "cabi.c"
    src.write_text('#include "b2g.h"\nint main(void) { return b2g_last_error() == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-fsyntax-only", str(src)])
    header = open(os.path.join(ROOT, "include", "b2g.h")).read()
    assert "#include <torch" not in header and "cuda_runtime" not in header and "at::" not in header


LINK_PROPS = ["mass", "ground_friction", "ground_torque_integral", "contact_friction", "inertia", "com_inertia",
              "ground_friction_limit_force", "ground_friction_limit_torque"]


def test_link_properties_match_oracle(b2g, oracle):
    """Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction
    (Box2DWorld.cpp:478-547) applied to the model: mass data, ground-friction joint limits, fixture friction and every getter
    agree bit for bit with the oracle's restatement (b2Body::Get/SetMassData, updateFrictionJoint), incl. static links."""
    m = b2g.Model(b2g.data_path(ENV_A))
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    names = {o["name"]: o for o in m.objects}
    robot_base, obj1, obj2 = m.objects[0]["first_body"], names["obj1"]["first_body"], names["obj2"]["first_body"]
    assert m.objects[0]["is_robot"] and names["obj2"]["is_static"]
    edits = [(obj1, "mass", 0.37), (robot_base, "mass", 2.5), (robot_base + 2, "mass", 0.011), (obj2, "mass", 3.0),
             (obj1, "ground_friction", 0.31), (robot_base + 1, "ground_friction", 0.0), (obj1, "ground_torque_integral", 0.07),
             (obj2, "ground_friction", 0.9), (obj1, "contact_friction", 1.25), (robot_base, "contact_friction", 0.05),
             (obj1, "mass", 0.5)]
    for body, prop, value in edits:
        m.set_link_property(body, prop, value)
        w.set_link_property(body, prop, value)
        assert np.array_equal(m.body_model(), w.body_model()), (body, prop)
        assert np.array_equal(m.fixture_model()[1], w.fixture_model()[1]), (body, prop)
        assert np.array_equal(m.joint_model()[1][:, 6:], w.joints()[1][:, 6:]), (body, prop)
        for b in range(1, m.n_bodies):
            for p in LINK_PROPS:
                assert m.get_link_property(b, p) == w.get_link_property(b, p), (b, p)
    with pytest.raises(b2g.B2GError):
        m.set_link_property(0, "mass", 1.0)               # the ground is not a link
    with pytest.raises(b2g.B2GError):
        m.set_link_property(obj1, "inertia", 1.0)         # derived, read-only
    for prop, value in [("mass", float("nan")), ("ground_friction", -0.1), ("contact_friction", float("inf"))]:
        with pytest.raises(b2g.B2GError):
            m.set_link_property(obj1, prop, value)        # upstream asserts on these
    assert np.array_equal(m.body_model(), w.body_model())   # refused calls leave the model alone


def test_link_property_known_answers(b2g):
    """setMass scales mass and inertia by the same ratio (Box2DWorld.cpp:483-487); the ground-friction joint follows with
    maxForce = mu m g scale, maxTorque = mu integral scale^2 (:529-530); the limit getters are unscaled (:502-510)."""
    m = b2g.Model(b2g.data_path(ENV_A))
    obj1 = {o["name"]: o for o in m.objects}["obj1"]["first_body"]
    mass0, inertia0 = m.get_link_property(obj1, "mass"), m.get_link_property(obj1, "inertia")
    m.set_link_property(obj1, "mass", 3.0 * mass0)
    m.set_link_property(obj1, "ground_friction", 0.25)
    m.set_link_property(obj1, "ground_torque_integral", 0.02)
    assert m.get_link_property(obj1, "mass") == pytest.approx(3.0 * mass0, rel=1e-6)
    assert m.get_link_property(obj1, "inertia") == pytest.approx(3.0 * inertia0, rel=1e-5)
    assert m.body_model()[obj1, 1] == pytest.approx(1.0 / (3.0 * mass0), rel=1e-6)
    ji, jf = m.joint_model()
    fj = [j for j in range(m.n_joints) if ji[j, 0] == 0 and ji[j, 2] == obj1][0]
    assert jf[fj, 6] == pytest.approx(0.25 * 3.0 * mass0 * 9.81 * m.scale, rel=1e-6)
    assert jf[fj, 7] == pytest.approx(0.25 * 0.02 * m.scale ** 2, rel=1e-6)
    assert m.get_link_property(obj1, "ground_friction_limit_force") == pytest.approx(0.25 * 3.0 * mass0 * 9.81, rel=1e-6)
    assert m.get_link_property(obj1, "ground_friction_limit_torque") == pytest.approx(0.25 * 0.02, rel=1e-6)


def test_link_and_joint_lookup_by_name(b2g):
    """Box2DObject::getLink(name) / getJoint(name) (Box2DWorld.cpp:1411-1505): names <-> body / joint indices, checked
    against the YAML (test_robot.yaml: base + 2 x 2 finger links, 4 actuated limited joints in DOF order 3..6)."""
    m = b2g.Model(b2g.data_path(ENV_A))
    env = load_env_dict(ENV_A)
    robot = env["robots"][0]
    for k, ld in enumerate(robot["links"]):
        body = m.find_link(0, ld["name"])
        info = m.link(body)
        assert info["name"] == ld["name"] and info["object"] == 0 and not info["is_static"]
        assert info["is_base_link"] == (ld["name"] == robot.get("base_link", robot["links"][0]["name"]))
        assert info["n_fixtures"] == len(ld["geometry"])
    for k, jd in enumerate(robot["joints"]):
        j = m.find_joint(0, jd["name"])
        info = m.joint(j)
        assert info["name"] == jd["name"] and info["kind"] == 1 and info["dof_index"] == 3 + k
        assert info["body_a"] == m.find_link(0, jd["link_a"]) and info["body_b"] == m.find_link(0, jd["link_b"])
        assert info["actuated"] == bool(jd["actuated"]) and info["limited"]
        assert m.link(info["body_b"])["parent_joint"] == j
    obj2 = m.object_index("obj2")
    assert m.link(m.find_link("obj2", "test_object_link_2"))["is_static"]
    assert m.joint(m.find_joint(obj2, "object_joint"))["dof_index"] == m.objects[obj2]["dof_offset"]
    friction = [m.joint(j) for j in range(m.n_joints) if m.joint(j)["kind"] == 0]
    assert len(friction) == m.n_bodies - 1 and all(f["name"] == "" and f["body_a"] == 0 and f["dof_index"] == -1 for f in friction)
    for bad in [lambda: m.find_link(0, "nope"), lambda: m.find_joint(0, "object_joint"), lambda: m.link(0), lambda: m.joint(99),
                lambda: m.find_link(7, "x")]:
        with pytest.raises(b2g.B2GError):
            bad()


@pytest.mark.parametrize("env_name", [ENV_A, ENV_B])
def test_local_bounding_boxes(b2g, oracle, env_name):
    """Box2DLink::getLocalBoundingBox (Box2DWorld.cpp:69-90, 258-261) and Box2DObject::getLocalAABB (:1304-1312, 1928-1931)
    against the oracle's restatement and the YAML: obj1 is a 0.1 m square about its centre of mass; obj2's two links merge to
    [-0.05, 0.4] x [-0.05, 0.1] (a static object keeps the box as given: no centre of mass to shift by)."""
    m = b2g.Model(b2g.data_path(env_name))
    w = oracle.OracleWorld(load_env_dict(env_name))
    for b in range(1, m.n_bodies):
        assert np.array_equal(m.link_aabb(b), w.link_aabb(b))
    for o in range(m.n_objects):
        assert np.array_equal(m.object_aabb(o), w.object_aabb(o))
    assert np.allclose(m.object_aabb("obj1"), [-0.05, -0.05, 0.05, 0.05], atol=1e-6)
    assert np.allclose(m.object_aabb("obj2"), [-0.05, -0.05, 0.4, 0.1], atol=1e-6)
    robot = m.object_aabb(0)
    base = m.objects[0]["first_body"]
    lc = m.body_model()[base, 4:6] / m.scale                      # the robot's box is shifted by the base link's centre of mass
    merged = np.array([min(m.link_aabb(b)[0] for b in range(base, base + m.objects[0]["n_links"])),
                       min(m.link_aabb(b)[1] for b in range(base, base + m.objects[0]["n_links"])),
                       max(m.link_aabb(b)[2] for b in range(base, base + m.objects[0]["n_links"])),
                       max(m.link_aabb(b)[3] for b in range(base, base + m.objects[0]["n_links"]))], np.float32)
    assert np.allclose(robot, merged - np.array([lc[0], lc[1], lc[0], lc[1]], np.float32), atol=1e-6)
    with pytest.raises(b2g.B2GError):
        m.link_aabb(0)
    with pytest.raises(b2g.B2GError):
        m.object_aabb(17)


def test_non_finite_numbers_are_refused(b2g, tmp_path):
    """A `.inf` / `.nan` / overflowing literal anywhere in the description would poison every world after one step: the model
    builder names the place and refuses (found by mutating the shipped YAML files, 3000 variants, no crash)."""
    env = load_env_dict(ENV_A)
    for mutate, where in [(lambda e: e.__setitem__("world_bounds", [-20.0, float("-inf"), 20.0, 20.0]), "world_bounds"),
                          (lambda e: e["robots"][0]["links"][1].__setitem__("mass", float("nan")), "link"),
                          (lambda e: e["objects"][0]["links"][0]["geometry"][0].__setitem__(3, float("inf")), "geometry"),
                          (lambda e: e["robots"][0]["joints"][0].__setitem__("axis", [float("nan"), 0.1]), "joint"),
                          (lambda e: e["states"]["obj1"]["velocity"].__setitem__(0, float("inf")), "state")]:
        import copy
        bad = copy.deepcopy(env)
        mutate(bad)
        with pytest.raises(b2g.B2GError) as ei:
            b2g.Model(bad)
        assert "non-finite" in str(ei.value) and where in str(ei.value)
    for key in ["mass", "ground_friction", "ground_torque_integral", "contact_friction", "restitution"]:
        bad = copy.deepcopy(env)
        bad["objects"][0]["links"][0][key] = -0.1
        with pytest.raises(b2g.B2GError) as ei:
            b2g.Model(bad)                                     # upstream asserts on these
        assert "negative" in str(ei.value)
    text = open(b2g.data_path(ENV_A)).read().replace("world_bounds: [-20.0,", "world_bounds: [-20.1e9990,")
    assert "1e9990" in text
    import shutil
    data_dir = os.path.dirname(b2g.data_path(ENV_A))
    shutil.copytree(data_dir, str(tmp_path / "data"))
    (tmp_path / "data" / ENV_A).write_text(text)
    with pytest.raises(b2g.B2GError):
        b2g.Model(str(tmp_path / "data" / ENV_A))
