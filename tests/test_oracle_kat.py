"""Known answers derived from first principles pin the oracle's shape, mass and ordering
arithmetic.  The reference itself holds no golden vectors for this path (parity unpinned), so these double-precision
derivations are the anchor."""
import numpy as np
import pytest

from parity_utils import ENV_A, ENV_B, load_env_dict


def test_sincos_accuracy(oracle):
    x = np.concatenate([np.linspace(-40, 40, 200001), np.array([0.0, 1e-8, -1e-8, np.pi / 2, np.pi, 100.0, -250.0])]).astype(np.float32)
    s, c = oracle.sincos(x)
    es = np.abs(s.astype(np.float64) - np.sin(x.astype(np.float64)))
    ec = np.abs(c.astype(np.float64) - np.cos(x.astype(np.float64)))
    assert es.max() < 2.5e-7 and ec.max() < 2.5e-7     # <= ~2 ulp at magnitude 1
    s0, c0 = oracle.sincos(np.zeros(1, np.float32))
    assert s0[0] == 0.0 and c0[0] == 1.0


def test_hull_order_square(oracle):
    # obj1: 0.1 x 0.1 square scaled by 10 -> hull order [1,2,3,0]
    p = oracle.polygon_set([[0, 0], [1, 0], [1, 1], [0, 1]], density=0.1)
    assert p["count"] == 4
    assert p["vertices"].tolist() == [[1, 0], [1, 1], [0, 1], [0, 0]]
    assert p["normals"].tolist() == [[1, 0], [0, 1], [-1, 0], [0, -1]]
    assert p["mass"] == pytest.approx(0.1, rel=1e-6)
    assert p["center"].tolist() == pytest.approx([0.5, 0.5], rel=1e-6)
    # inertia about the body origin: I_com + m*|c|^2 = 0.1/6 + 0.1*0.5
    assert p["I"] == pytest.approx(0.1 / 6 + 0.05, rel=1e-5)


def test_hull_order_triangle(oracle):
    # robot base triangle [(0,0),(4,-4),(8,0)] -> [2,0,1]
    p = oracle.polygon_set([[0, 0], [4, -4], [8, 0]])
    assert p["vertices"].tolist() == [[8, 0], [0, 0], [4, -4]]


def test_hull_welds_and_drops_collinear(oracle):
    p = oracle.polygon_set([[0, 0], [1, 0], [2, 0], [2, 2], [0, 2], [0.00001, 0.00001]])
    assert p["count"] == 4


def test_mass_data_shape_a(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    bm = w.body_model()
    # mass, local COM, I about COM, invMass, invI
    expect = {
        1: (10.1, (4.0, -0.1666667), 60.3194, 0.0990099, 0.0165784),
        2: (0.3, (0.0, 2.0), 0.425, 3.33333, 2.35294),
        3: (0.1, (0.0, 1.0), 0.0416667, 10.0, 24.0),
        4: (0.3, (0.0, 2.0), 0.425, 3.33333, 2.35294),
        5: (0.1, (0.0, 1.0), 0.0416667, 10.0, 24.0),
        6: (0.1, (0.5, 0.5), 0.0166667, 10.0, 60.0),
    }
    for b, (m, com, I, im, ii) in expect.items():
        assert bm[b, 0] == pytest.approx(m, rel=2e-6)
        assert bm[b, 4:6].tolist() == pytest.approx(com, rel=2e-6, abs=1e-6)
        assert bm[b, 2] == pytest.approx(I, rel=5e-6)
        assert bm[b, 1] == pytest.approx(im, rel=5e-6)
        assert bm[b, 3] == pytest.approx(ii, rel=5e-6)
    for b in (0, 7, 8):   # ground and the static obj2 links carry no mass in Box2D
        assert bm[b, :4].tolist() == [0, 0, 0, 0]


def test_mass_data_rigid_base(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_B))
    bm = w.body_model()
    assert bm[1, 0] == pytest.approx(10.1, rel=2e-6)
    assert bm[1, 4:6].tolist() == pytest.approx([4.0, 0.468254], rel=2e-6)
    assert bm[1, 2] == pytest.approx(85.8398, rel=5e-6)
    assert bm[1, 3] == pytest.approx(0.0116496, rel=5e-6)


def test_creation_order_and_joints(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    assert (w.n_bodies, w.n_proxies, w.n_joints, w.nq) == (9, 10, 13, 11)
    jio, jfo = w.joints()
    # fric_base, fric_fl1, fric_fl2, fric_fr1, fric_fr2, rev_l1, rev_l2, rev_r1, rev_r2, fric_obj1, fric_o2a, fric_o2b, rev_obj2
    assert jio[:, 0].tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1]
    assert jio[:, 1].tolist() == [0, 0, 0, 0, 0, 1, 2, 1, 4, 0, 0, 0, 7]
    assert jio[:, 2].tolist() == [1, 2, 3, 4, 5, 2, 3, 4, 5, 6, 7, 8, 8]
    # revolute anchors in the parent frame (Box2D units) and limits
    assert jfo[5, 8:10].tolist() == pytest.approx([1, 1]) and jfo[6, 8:10].tolist() == pytest.approx([0, 3.5])
    assert jfo[7, 8:10].tolist() == pytest.approx([7, 1]) and jfo[8, 8:10].tolist() == pytest.approx([0, 3.5])
    assert np.allclose(jfo[5:9, 13], -1.05) and np.allclose(jfo[5:9, 14], 1.05)
    # friction joints: maxForce = mu_g * 98.1 * m, maxTorque = mu_g * integral * 100
    assert jfo[0, 6] == pytest.approx(0.1 * 98.1 * 10.1, rel=1e-5) and jfo[0, 7] == pytest.approx(0.1 * 0.1 * 100, rel=1e-5)
    assert jfo[9, 6] == pytest.approx(0.001 * 98.1 * 0.1, rel=1e-5)


def test_initial_state_matches_yaml(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    q, qd = w.get_state()
    assert q.tolist() == pytest.approx([0, 0, 0, 0, 0, 0, 0, 0.9, 1.0, -1.0, 0.0], abs=2e-7)
    assert qd.tolist() == pytest.approx([0, 0, 0, 0, -5.5, 0, 0, -0.2, -0.1, 1.0, 0.0], abs=1e-6)
    # initial obj1 transform the reference's test passes to the sim_env suite (tests/test_Box2DWorld.cpp:38)
    assert w.get_object_pose(1).tolist() == pytest.approx([0.9, 1.0, -1.0], abs=2e-7)
    # static obj2 at (2.9, 1.2, -6.12) (test_env.yaml:23-25)
    assert w.get_object_pose(2).tolist() == pytest.approx([2.9, 1.2, -6.12], abs=1e-6)


def test_island_order_no_contacts(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    w.trace_islands(True)
    w.step(1)
    tr = w.island_trace()
    # island 1 seeds at obj1, island 2 at finger_r_2; joints in DFS order
    assert tr[0]["bodies"] == [6, 0] and tr[0]["joints"] == [9]
    assert tr[1]["bodies"] == [5, 0, 4, 1, 2, 3]
    assert tr[1]["joints"] == [8, 4, 7, 3, 5, 0, 6, 1, 2]


def test_momentum_free_flight(oracle):
    """With (almost) no ground friction and no contacts a box keeps its velocity: checks the integrator."""
    env = load_env_dict(ENV_A)
    for o in env["objects"]:
        for l in o["links"]:
            l["ground_friction"] = 0.0
    env["states"]["obj1"] = dict(configuration=[-1.5, -1.5, 0.3], velocity=[0.3, -0.2, 0.7])
    w = oracle.OracleWorld(env)
    q0, qd0 = w.get_state()
    w.step(50)
    q1, qd1 = w.get_state()
    assert qd1[7:10].tolist() == pytest.approx([0.3, -0.2, 0.7], rel=1e-6)
    assert q1[7:10].tolist() == pytest.approx([-1.5 + 0.5 * 0.3, -1.5 - 0.5 * 0.2, 0.3 + 0.5 * 0.7], rel=1e-5)


def test_ground_friction_decelerates(oracle):
    """Coulomb ground friction: a sliding box loses mu*g per second (b2FrictionJoint maxForce = mu*m*g)."""
    env = load_env_dict(ENV_A)
    for l in env["objects"][0]["links"]:
        l["ground_friction"] = 0.05
        l["ground_torque_integral"] = 0.0
    env["states"]["obj1"] = dict(configuration=[-1.5, -1.5, 0.0], velocity=[1.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    w.step(100)
    _, qd = w.get_state()
    assert qd[7] == pytest.approx(1.0 - 0.05 * 9.81 * 1.0, rel=2e-3)


def test_joint_limits_respected_and_anchor_drift(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    w.set_target_velocity(0, [0.5, 0.0, 0.3, 0.5, -0.5, -0.5, 0.5])
    for _ in range(300):
        w.step(1)
        q, _ = w.get_state()
        assert np.all(np.abs(q[3:7]) <= 1.05 + 2.0 / 180.0 * np.pi + 1e-3)
    b = w.bodies()
    # revolute anchor drift <= linearSlop-ish: child origin (xf.p) vs parent anchor
    jio, jfo = w.joints()
    for j in (5, 6, 7, 8):
        a, c = jio[j, 1], jio[j, 2]
        s, co = b[a, 2], b[a, 3]
        ax = co * jfo[j, 8] - s * jfo[j, 9] + b[a, 0]
        ay = s * jfo[j, 8] + co * jfo[j, 9] + b[a, 1]
        assert np.hypot(ax - b[c, 0], ay - b[c, 1]) < 0.02


def test_velocity_controller_reaches_target(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    w.set_target_velocity(0, [0.3, -0.2, 0.2, 0, 0, 0, 0])
    w.step(200)
    _, qd = w.get_state()
    assert qd[:3].tolist() == pytest.approx([0.3, -0.2, 0.2], abs=0.05)


def test_sleep_zeroes_velocity(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    q, qd = w.get_state()
    qd[:] = 0
    w.set_state(q, qd)
    w.step(80)   # > timeToSleep / dt = 50 steps of rest
    b = w.bodies()
    assert np.all(b[1:7, 14] == 0.0)        # every dynamic body asleep
    assert np.all(b[1:7, 10:13] == 0.0)


def test_contact_and_toi_vs_static(oracle):
    """A box thrown at the static obstacle must not tunnel (TOI path) and must bounce (restitution 1)."""
    env = load_env_dict(ENV_A)
    env["states"]["obj1"] = dict(configuration=[2.9, 0.4, 0.0], velocity=[0.0, 8.0, 0.0])
    w = oracle.OracleWorld(env)
    toi_events = 0
    for _ in range(40):
        w.step(1)
        toi_events += w.stats()["toi_events"]
    q, qd = w.get_state()
    assert toi_events > 0
    assert q[8] < 1.2            # did not pass through the bar centred at y = 1.2
    assert qd[8] < 0.0           # bounced back


def test_region_query_known_answers(oracle):
    """getObjects(aabb): obj1 sits at (0.9, 1.0) in test_env.yaml, the robot at the origin, obj2 at (2.9, 1.2)."""
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    assert w.objects_in_aabb([0.5, 0.5, 1.5, 1.5], exclude_robots=False).tolist() == [False, True, False]
    assert w.objects_in_aabb([-5, -5, 5, 5], exclude_robots=True).tolist() == [False, True, True]
    assert w.objects_in_aabb([-5, -5, 5, 5], exclude_robots=False).tolist() == [True, True, True]
    assert not w.objects_in_aabb([10, 10, 11, 11], exclude_robots=False).any()
    # a box that only reaches the fat AABB margin of obj1 but not its shape must not report it (exact GJK test)
    lo_x = 0.9 + 0.1 + 0.0015 + 0.003   # obj1 is a 0.2 m box: half width + polygon radius (1 mm) ... + a bit
    assert not w.objects_in_aabb([lo_x + 0.05, 0.0, lo_x + 0.5, 2.0], exclude_robots=False)[1]


def test_feasibility_known_answers(oracle):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    assert w.is_physically_feasible()
    w.set_object_pose(1, 0.2, 0.0, 0.0)      # obj1 pushed into the robot base: overlap area >> 1e-4 m^2
    assert not w.is_physically_feasible()
    w.set_object_pose(1, 3.0, -3.0, 0.0)     # far away again
    assert w.is_physically_feasible()


# configurations / velocities the reference's own test hands to the sim_env suite (tests/test_Box2DWorld.cpp:11-20, 46-51,
# 67-72): a set followed by a get has to return them
REF_OBJECT = (1, [1.0, 2.0, 3.0], [0.0, 1.0, 0.2])
REF_ROBOT = (0, [0.0, 6.0, 2.0, 0.1, 0.1, 0.1, 0.1], [0.0, 1.0, 0.2, 0.1, 0.2, 0.0, 0.1])


@pytest.mark.parametrize("obj,conf,vel", [REF_OBJECT, REF_ROBOT])
def test_reference_test_configurations_round_trip(oracle, obj, conf, vel):
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    w.set_object_dofs(obj, conf)
    w.set_object_dofs(obj, vel, velocities=True)
    assert w.get_object_dofs(obj).tolist() == pytest.approx(conf, abs=2e-6)
    assert w.get_object_dofs(obj, velocities=True).tolist() == pytest.approx(vel, abs=2e-6)
    pose = w.get_object_pose(obj)
    assert pose.tolist() == pytest.approx(conf[:3], abs=2e-6)


def test_collide_polygons_box_face_known_answers(oracle):
    """b2CollidePolygons on two axis-aligned unit boxes, derived by hand from the published algorithm (Box2D 2.3.1
    b2CollidePolygon.cpp) -- A at the origin, B shifted along +x:

    hull order of a box after b2PolygonShape::Set (KAT above): v0 = (h,-h), v1 = (h,h), v2 = (-h,h), v3 = (-h,-h), normals
    (1,0) (0,1) (-1,0) (0,-1).  b2FindMaxSeparation(A, B) = x_B - 2h on A's edge 0, b2FindMaxSeparation(B, A) the same on B's
    edge 2; equal separations -> no flip (k_tol = 0.1 * linearSlop), reference face = A's edge 0, manifold type e_faceA with
    localNormal (1,0), localPoint (h,0).  Incident edge on B = the edge whose normal is most anti-parallel: edge 2, vertices
    v2, v3 -> clip points (x_B - h, +h) and (x_B - h, -h), both inside the side planes (|y| <= h + totalRadius), kept while
    separation = x_B - 2h <= totalRadius = 2 * polygonRadius = 0.02; manifold points are stored in B's frame: (-h, +h),
    (-h, -h); contact ids {indexA = 0 (reference edge), indexB = 2 / 3 (incident vertex), typeA = face (1), typeB = vertex
    (0)} = 0x010200, 0x010300."""
    h = 0.5
    box = [-h, -h, h, -h, h, h, -h, h]
    for xb, expect_points in ((0.9, 2), (1.015, 2), (1.03, 0)):     # overlapping / inside the 0.02 skin / separated
        m = oracle.collide_polygons(box, [0, 0, 0], box, [xb, 0, 0])
        assert m["point_count"] == expect_points
        if expect_points == 0:
            continue
        assert m["type"] == 1                                            # b2Manifold::e_faceA
        assert m["local_normal"].tolist() == [1.0, 0.0]
        assert m["local_point"].tolist() == [0.5, 0.0]
        assert np.allclose(m["points"], [[-0.5, 0.5], [-0.5, -0.5]], atol=1e-6)
        assert m["ids"] == [0x010200, 0x010300]
    # the mirrored configuration makes A's edge 2 the reference face and B's vertices v0, v1 the incident ones
    m = oracle.collide_polygons(box, [0, 0, 0], box, [-0.9, 0, 0])
    assert (m["point_count"], m["type"]) == (2, 1)
    assert m["local_normal"].tolist() == [-1.0, 0.0] and m["local_point"].tolist() == [-0.5, 0.0]
    assert np.allclose(m["points"], [[0.5, -0.5], [0.5, 0.5]], atol=1e-6)
    assert m["ids"] == [0x010002, 0x010102]


def _two_box_env(restitution):
    def box(name):
        link = dict(name="link", geometry=[[-0.1, -0.1, 0.1, -0.1, 0.1, 0.1, -0.1, 0.1]], mass=0.5, ground_friction=0.0,
                    ground_torque_integral=0.0, contact_friction=0.0, restitution=restitution, balls=[])
        return dict(name=name, static=False, links=[link], joints=[], base_actuation=None)
    return dict(scale=10.0, world_bounds=[-5, -5, 5, 5], robots=[], objects=[box("a"), box("b")],
                states={"a": dict(configuration=[0.0, 0.0, 0.0], velocity=[1.0, 0.0, 0.0]),
                        "b": dict(configuration=[0.5, 0.0, 0.0], velocity=[0.0, 0.0, 0.0])})


@pytest.mark.parametrize("restitution,va,vb", [(1.0, 0.0, 1.0), (0.0, 0.5, 0.5)])
def test_head_on_collision_known_answers(oracle, restitution, va, vb):
    """Two equal frictionless boxes, a at 1 m/s onto b at rest (no ground friction, no gravity): conservation of momentum
    plus Newton's restitution law give the textbook answers -- velocities swap for e = 1 (b2MixRestitution = max, relative
    normal speed 10 > b2_velocityThreshold 1 in Box2D units), both move at 0.5 m/s for e = 0."""
    w = oracle.OracleWorld(_two_box_env(restitution))
    w.step(60)
    q, qd = w.get_state()
    assert qd[0] + qd[3] == pytest.approx(1.0, abs=1e-5)              # momentum (equal masses)
    assert qd[0] == pytest.approx(va, abs=2e-3) and qd[3] == pytest.approx(vb, abs=2e-3)
    assert abs(qd[1]) < 1e-6 and abs(qd[4]) < 1e-6 and abs(qd[2]) < 1e-5 and abs(qd[5]) < 1e-5   # stays one-dimensional
    if restitution == 0.0:
        assert q[3] - q[0] >= 0.2 - 0.011                             # boxes rest against each other, within the slop


def test_distance_known_answers(oracle):
    """b2Distance (GJK) between boxes, from plane geometry: face to face -> the gap; corner to corner -> the diagonal of the
    gap; a box rotated by 45 degrees -> centre distance minus the half diagonal minus the half width; with radii the
    two polygon radii (0.01 each) come off; overlapping shapes -> 0."""
    h = 0.5
    box = [-h, -h, h, -h, h, h, -h, h]
    d = oracle.distance(box, [0, 0, 0], box, [3.0, 0.2, 0])
    assert d["distance"] == pytest.approx(2.0, abs=1e-6) and d["point_a"][0] == pytest.approx(0.5, abs=1e-6)
    assert oracle.distance(box, [0, 0, 0], box, [3.0, 0.2, 0], use_radii=True)["distance"] == pytest.approx(1.98, abs=1e-6)
    d = oracle.distance(box, [0, 0, 0], box, [2.0, 3.0, 0])
    assert d["distance"] == pytest.approx(np.hypot(1.0, 2.0), abs=1e-6)
    assert np.allclose(d["point_a"], [0.5, 0.5], atol=1e-6) and np.allclose(d["point_b"], [1.5, 2.5], atol=1e-6)
    d = oracle.distance(box, [0, 0, 0], box, [3.0, 0.0, np.pi / 4])
    assert d["distance"] == pytest.approx(3.0 - 0.5 - np.sqrt(0.5), abs=2e-6)
    assert oracle.distance(box, [0, 0, 0], box, [0.7, 0.1, 0.3])["distance"] == pytest.approx(0.0, abs=1e-6)


def test_time_of_impact_known_answers(oracle):
    """b2TimeOfImpact for a box translating onto a resting box: contact is declared when the separation reaches the target
    = max(linearSlop, totalRadius - 3 linearSlop) = 0.005 (totalRadius = 2 * 0.01), within a tolerance of 0.25 * linearSlop
    -> t = (gap - target) / travelled distance.  A sweep that stops short -> separated at t = 1."""
    h = 0.5
    box = [-h, -h, h, -h, h, h, -h, h]
    # gap 2.0, B travels 4.0 towards A: t = (2.0 - 0.005) / 4.0
    state, t = oracle.time_of_impact(box, [0, 0, 0], [0, 0, 0], box, [3.0, 0.1, 0], [-1.0, 0.1, 0])
    assert state == 3 and t == pytest.approx((2.0 - 0.005) / 4.0, abs=0.25 * 0.005 / 4.0 + 1e-6)
    # both moving: closing distance 1.5 + 2.5
    state, t = oracle.time_of_impact(box, [0, 0, 0], [1.5, 0, 0], box, [3.0, 0.1, 0], [0.5, 0.1, 0])
    assert state == 3 and t == pytest.approx((2.0 - 0.005) / 4.0, abs=0.25 * 0.005 / 4.0 + 1e-6)
    # stops one unit short
    state, t = oracle.time_of_impact(box, [0, 0, 0], [0, 0, 0], box, [3.0, 0.1, 0], [2.0, 0.1, 0])
    assert state == 4 and t == 1.0


def test_hinged_pair_conserves_momentum(oracle):
    """Two bars joined by a free hinge, no ground friction, no contacts, no gravity: the revolute joint exchanges equal and
    opposite impulses at one point, so total linear momentum and total angular momentum (about the origin) are conserved
    while the relative angle changes and the anchors stay together."""
    def link(name, poly, mass):
        return dict(name=name, geometry=[poly], mass=mass, ground_friction=0.0, ground_torque_integral=0.0, contact_friction=0.0,
                    restitution=0.0, balls=[])
    obj = dict(name="pair", static=False, base_actuation=None,
               links=[link("a", [-0.2, -0.05, 0.2, -0.05, 0.2, 0.05, -0.2, 0.05], 0.7), link("b", [0.0, -0.04, 0.3, -0.04, 0.3, 0.04, 0.0, 0.04], 0.4)],
               joints=[dict(name="hinge", link_a="a", link_b="b", axis=[0.2, 0.0], axis_orientation=0.0, position_limits=[0.0, 0.0],
                            velocity_limits=[-100.0, 100.0], acceleration_limits=[-100.0, 100.0], joint_type="revolute", actuated=False)])
    env = dict(scale=10.0, world_bounds=[-5, -5, 5, 5], robots=[], objects=[obj],
               states={"pair": dict(configuration=[0.1, -0.2, 0.3, 0.4], velocity=[0.3, -0.2, 0.8, -1.5])})
    w = oracle.OracleWorld(env)
    bm = w.body_model()            # mass invMass I (about the centre of mass, b2Body::m_I) invI lcx lcy type nfix (Box2D units)

    def momenta():
        b = w.bodies()
        p = np.zeros(2)
        L = 0.0
        for i in (1, 2):
            m, I_c = float(bm[i, 0]), float(bm[i, 2])
            c, v, om = b[i, 4:6].astype(np.float64), b[i, 10:12].astype(np.float64), float(b[i, 12])
            p += m * v
            L += I_c * om + m * (c[0] * v[1] - c[1] * v[0])
        return p, L
    p0, L0 = momenta()
    q0, _ = w.get_state()
    w.step(200, allow_sleeping=False)
    p1, L1 = momenta()
    q1, _ = w.get_state()
    assert np.allclose(p1, p0, rtol=0, atol=1e-4 * np.abs(p0).max())
    assert L1 == pytest.approx(L0, rel=2e-3)
    assert abs(q1[3] - q0[3]) > 0.5                           # the hinge really turned
    b = w.bodies()                                            # anchors coincide within the linear slop (Box2D units)
    ca, sa = np.cos(b[1, 6]), np.sin(b[1, 6])
    anchor_a = b[1, 0:2] + np.array([ca * 2.0, sa * 2.0])     # body a origin + R * (scale * axis)
    assert np.linalg.norm(anchor_a - b[2, 0:2]) < 0.02


def test_step_translation_and_rotation_clamps(oracle):
    """b2Island::Solve clamps the motion of one step to b2_maxTranslation = 2 and b2_maxRotation = pi/2 (Box2D units) by
    scaling the velocity: with scale 10 and dt 0.01 a box thrown at 50 m/s covers 0.2 m in the step and leaves it at
    20 m/s; one spun at 400 rad/s turns a quarter turn and leaves at 50 pi rad/s."""
    env = _two_box_env(0.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[50.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.0, 3.0, 0.0], velocity=[0.0, 0.0, 400.0])
    w = oracle.OracleWorld(env)
    w.step(1)
    q, qd = w.get_state()
    assert q[0] == pytest.approx(0.2, rel=1e-6) and qd[0] == pytest.approx(20.0, rel=1e-6)
    assert q[5] == pytest.approx(np.pi / 2, rel=1e-6) and qd[5] == pytest.approx(50.0 * np.pi, rel=1e-6)
    assert q[1] == 0.0 and q[3] == pytest.approx(0.0, abs=1e-6) and q[4] == pytest.approx(3.0, abs=1e-6)


def test_restitution_needs_the_velocity_threshold(oracle):
    """b2ContactSolver only applies restitution when the approach speed exceeds b2_velocityThreshold = 1 (Box2D units):
    at 0.05 m/s (0.5 units) two fully elastic boxes still collide inelastically and share the momentum."""
    env = _two_box_env(1.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[0.05, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.25, 0.0, 0.0], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    w.step(200)
    q, qd = w.get_state()
    assert qd[0] + qd[3] == pytest.approx(0.05, abs=1e-6)
    assert qd[0] == pytest.approx(0.025, abs=5e-4) and qd[3] == pytest.approx(0.025, abs=5e-4)


def test_overlap_recovers_to_the_linear_slop(oracle):
    """Two boxes created 1 cm inside each other: the position solver (pseudo impulses, Baumgarte 0.2, no velocity added)
    pushes them apart until the separation of the skinned polygons is -b2_linearSlop, i.e. the faces end
    2 * b2_polygonRadius - b2_linearSlop = 0.015 Box2D units = 1.5 mm apart, symmetrically and at rest."""
    env = _two_box_env(0.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[0.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.19, 0.0, 0.0], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    w.step(1)
    q, _ = w.get_state()
    assert -0.01 < q[3] - q[0] - 0.2 < 0.0015          # on its way after one step, never past the target
    w.step(150)
    q, qd = w.get_state()
    assert q[3] - q[0] - 0.2 == pytest.approx(0.0015, abs=2e-6)
    assert q[0] + q[3] == pytest.approx(0.19, abs=1e-6) and np.all(qd == 0.0)
    # the two manifold points are corrected one after the other, which breaks the mirror symmetry by a few microns
    assert abs(q[1]) < 5e-5 and abs(q[2]) < 5e-4 and abs(q[4]) < 5e-5 and abs(q[5]) < 5e-4


def test_link_states_known_answers(oracle):
    """Box2DLink::getPose / getVelocityVector / getCenterOfMass: a dynamic object's base link has its origin at the centre of
    mass (setOriginToCenterOfMass), so pose == centre of mass == the object's base DOFs; a finger link sits at its parent's
    joint anchor (test_robot.yaml axis [-0.3, 0.2] relative to the base link origin before the origin moved to the COM)."""
    w = oracle.OracleWorld(load_env_dict(ENV_A))
    q, qd = w.get_state()
    ls = w.link_states()
    assert not ls[0].any()                                                    # the ground is not a link
    assert np.allclose(ls[1, :3], q[0:3], atol=1e-6) and np.allclose(ls[1, 6:8], q[0:2], atol=1e-6)
    assert np.allclose(ls[6, :3], q[7:10], atol=1e-6) and np.allclose(ls[6, 6:8], q[7:9], atol=1e-6)
    assert np.allclose(ls[6, 3:6], qd[7:10], atol=1e-6)
    assert np.allclose(ls[2, 0], -0.3, atol=1e-6) and np.allclose(ls[4, 0], 0.3, atol=1e-6)   # finger roots mirror each other
    assert ls[2, 1] == pytest.approx(ls[4, 1], abs=1e-6)
    assert np.allclose(ls[3, :2] - ls[2, :2], [0.0, 0.35], atol=1e-5)          # second finger joint 0.35 m up the first link
    # YAML initial velocity of finger_l_joint_2 (index 4): the outer link turns about its joint, its COM 0.1 m further out
    assert ls[3, 5] == pytest.approx(qd[4] + qd[3] + qd[2], abs=1e-5)
    assert ls[3, 3] == pytest.approx(-ls[3, 5] * (ls[3, 7] - ls[3, 1]), abs=1e-5)


def test_ground_friction_torque_decelerates_spin(oracle):
    """The angular row of b2FrictionJoint: a spinning box loses mu * torque_integral / I per second (maxTorque =
    mu * integral * scale^2 in Box2D units, Box2DWorld.cpp:145; inertia scales with scale^2 as well), and setMass /
    setGroundFrictionTorqueIntegral (:478-531) change that rate the way the formula says."""
    env = _two_box_env(0.0)
    link = env["objects"][0]["links"][0]
    link["ground_friction"], link["ground_torque_integral"] = 0.2, 0.004
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[0.0, 0.0, 3.0])
    env["states"]["b"] = dict(configuration=[0.0, 3.0, 0.0], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    inertia = w.get_link_property(1, "com_inertia")                     # 0.5 kg, 0.2 m square: m (w^2 + h^2) / 12
    assert inertia == pytest.approx(0.5 * (0.2 ** 2 + 0.2 ** 2) / 12.0, rel=1e-5)
    w.step(50)
    _, qd = w.get_state()
    rate = 0.2 * 0.004 / inertia
    assert qd[2] == pytest.approx(3.0 - rate * 0.5, rel=1e-3) and qd[0] == 0.0 and qd[1] == 0.0
    w.set_link_property(1, "mass", 1.0)                                 # twice the mass -> twice the inertia -> half the rate
    w.set_link_property(1, "ground_torque_integral", 0.002)             # and half the torque
    w0 = qd[2]
    w.step(50)
    _, qd = w.get_state()
    assert qd[2] == pytest.approx(w0 - 0.25 * rate * 0.5, rel=1e-3)


def test_joint_effort_is_a_constant_torque(oracle):
    """Box2DJoint::setControlTorque (Box2DWorld.cpp:1198-1222): motor speed +-FLT_MAX with maxMotorTorque = scale^2 |tau|
    makes the revolute motor saturate every step, i.e. it applies the constant torque tau.  With a base 10^4 times heavier
    than the arm, no ground friction and no limits, the arm spins up at tau / I_anchor (inertia of the arm about the hinge:
    m (w^2 + h^2) / 12 + m d^2, hinge at the arm's end) and the base takes the opposite angular momentum."""
    def link(name, poly, mass):
        return dict(name=name, geometry=[poly], mass=mass, ground_friction=0.0, ground_torque_integral=0.0, contact_friction=0.0,
                    restitution=0.0, balls=[])
    arm_m, arm_w, arm_h = 0.3, 0.1, 0.4
    robot = dict(name="r", static=False,
                 links=[link("base", [-0.5, -0.5, 0.5, -0.5, 0.5, 0.5, -0.5, 0.5], 3000.0),
                        link("arm", [-arm_w / 2, 0.0, arm_w / 2, 0.0, arm_w / 2, arm_h, -arm_w / 2, arm_h], arm_m)],
                 joints=[dict(name="hinge", link_a="base", link_b="arm", axis=[0.0, 0.6], axis_orientation=0.0, position_limits=[0.0, 0.0],
                              velocity_limits=[-100.0, 100.0], acceleration_limits=[-1000.0, 1000.0], joint_type="revolute",
                              actuated=True)],
                 base_actuation=dict(translational_velocity_limits=[-2.0, 2.0], translational_acceleration_limits=[-2.0, 2.0],
                                     rotational_velocity_limits=[-2.0, 2.0], rotational_acceleration_limits=[-2.0, 2.0],
                                     use_center_of_mass=True))
    env = dict(scale=10.0, world_bounds=[-5, -5, 5, 5], robots=[robot], objects=[],
               states={"r": dict(configuration=[0.0, 0.0, 0.0, 0.0], velocity=[0.0, 0.0, 0.0, 0.0])})
    w = oracle.OracleWorld(env)
    tau = 0.02
    w.set_control_efforts(0, np.array([0.0, 0.0, 0.0, tau], np.float32))
    w.step(40)
    q, qd = w.get_state()
    i_anchor = arm_m * (arm_w ** 2 + arm_h ** 2) / 12.0 + arm_m * (arm_h / 2) ** 2
    assert qd[3] == pytest.approx(tau / i_anchor * 0.4, rel=5e-3)                  # joint speed after 0.4 s
    assert q[3] == pytest.approx(0.5 * tau / i_anchor * 0.4 ** 2, rel=3e-2)        # semi-implicit Euler: within a step's worth
    assert abs(qd[2]) < 1e-3 * abs(qd[3])                                          # the heavy base hardly turns


def test_base_effort_is_a_force_on_the_robot(oracle):
    """Box2DRobot::commandEfforts (Box2DWorld.cpp:2319-2379): base efforts are a force scale * f at the base link's centre
    of mass and a torque scale^2 * tau.  Whatever the hinge does, Newton's second law for the whole robot holds: the total
    linear momentum after t seconds is f t (user units), and with a force through the base COM alone the free arm lags."""
    def link(name, poly, mass):
        return dict(name=name, geometry=[poly], mass=mass, ground_friction=0.0, ground_torque_integral=0.0, contact_friction=0.0,
                    restitution=0.0, balls=[])
    robot = dict(name="r", static=False,
                 links=[link("base", [-0.2, -0.2, 0.2, -0.2, 0.2, 0.2, -0.2, 0.2], 2.0),
                        link("arm", [-0.05, 0.0, 0.05, 0.0, 0.05, 0.4, -0.05, 0.4], 0.5)],
                 joints=[dict(name="hinge", link_a="base", link_b="arm", axis=[0.0, 0.3], axis_orientation=0.0, position_limits=[0.0, 0.0],
                              velocity_limits=[-100.0, 100.0], acceleration_limits=[-1000.0, 1000.0], joint_type="revolute",
                              actuated=True)],
                 base_actuation=dict(translational_velocity_limits=[-20.0, 20.0], translational_acceleration_limits=[-20.0, 20.0],
                                     rotational_velocity_limits=[-20.0, 20.0], rotational_acceleration_limits=[-20.0, 20.0],
                                     use_center_of_mass=True))
    env = dict(scale=10.0, world_bounds=[-5, -5, 5, 5], robots=[robot], objects=[],
               states={"r": dict(configuration=[0.0, 0.0, 0.0, 0.0], velocity=[0.0, 0.0, 0.0, 0.0])})
    w = oracle.OracleWorld(env)
    f = np.array([1.5, -0.5], np.float64)
    w.set_control_efforts(0, np.array([f[0], f[1], 0.0, 0.0], np.float32))
    w.step(30)
    bm, b = w.body_model(), w.bodies()
    p = sum(float(bm[i, 0]) * b[i, 10:12].astype(np.float64) for i in (1, 2)) / 10.0        # Box2D units -> user units
    assert np.allclose(p, f * 0.3, rtol=2e-4)
    ls = w.link_states()
    assert ls[1, 3] > ls[2, 3] > 0.0                       # pulled along through the hinge, the arm's COM lags behind the base
    assert np.allclose(ls[2, :2] - ls[1, :2], np.array([0.0, 0.3]) @ np.array([[np.cos(ls[1, 2]), np.sin(ls[1, 2])],
                                                                               [-np.sin(ls[1, 2]), np.cos(ls[1, 2])]]), atol=2e-3)


def test_velocity_controller_ramps_at_the_acceleration_limit(oracle):
    """Box2DRobotVelocityController::control (Box2DController.cpp:65-117): effort = mass * clamp((v* - v) / dt, accel limits)
    with mass = the robot's total mass (Box2DObject::_mass), applied as a force on the base link.  Far from the target the
    clamp is active, so without ground friction the robot's centre of mass accelerates at exactly the limit
    (test_robot.yaml: 2 m/s^2): total momentum = M a t."""
    env = load_env_dict(ENV_A)
    for l in env["robots"][0]["links"]:
        l["ground_friction"] = 0.0
    env["states"]["my_robot"]["velocity"] = [0.0] * 7
    env["states"]["obj1"] = dict(configuration=[3.0, 3.0, 0.0], velocity=[0.0, 0.0, 0.0])       # out of the way
    w = oracle.OracleWorld(env)
    w.set_target_velocity(0, np.array([2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], np.float32))
    w.step(40)
    bm, b = w.body_model(), w.bodies()
    links = range(1, 6)
    mass = sum(float(bm[i, 0]) for i in links)
    p = sum(float(bm[i, 0]) * b[i, 10:12].astype(np.float64) for i in links) / 10.0
    assert p[0] / mass == pytest.approx(2.0 * 0.4, rel=1e-3) and abs(p[1] / mass) < 1e-3
    w.step(100)                                                                                   # then it settles at the target
    _, qd = w.get_state()
    assert qd[0] == pytest.approx(2.0, abs=2e-2)


def test_off_centre_elastic_collision_conserves_everything(oracle):
    """A box hits a rotated box off-centre (corner against face: one manifold point), frictionless, restitution 1: both
    start to spin, and linear momentum, angular momentum about the origin and kinetic energy all survive the impact —
    normal-impulse-only contact with Newton restitution is exactly that."""
    env = _two_box_env(1.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[1.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.5, 0.13, 0.6], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    bm = w.body_model()

    def invariants():
        b = w.bodies()
        p, ang, energy = np.zeros(2), 0.0, 0.0
        for i in (1, 2):
            m, inertia = float(bm[i, 0]), float(bm[i, 2])
            c, v, om = b[i, 4:6].astype(np.float64), b[i, 10:12].astype(np.float64), float(b[i, 12])
            p += m * v
            ang += m * (c[0] * v[1] - c[1] * v[0]) + inertia * om
            energy += 0.5 * m * (v @ v) + 0.5 * inertia * om * om
        return p, ang, energy

    p0, l0, e0 = invariants()
    w.step(60)
    p1, l1, e1 = invariants()
    _, qd = w.get_state()
    assert abs(qd[2]) > 1.0 and abs(qd[5]) > 1.0 and abs(qd[1]) > 0.1           # the impact really was oblique
    assert np.allclose(p1, p0, atol=1e-5) and abs(l1 - l0) < 1e-5 and e1 == pytest.approx(e0, rel=1e-6)


@pytest.mark.parametrize("mu_a,mu_b", [(0.3, 0.3), (0.5, 0.2)])
def test_oblique_impact_obeys_coulomb_friction(oracle, mu_a, mu_b):
    """A box strikes the face of a resting box while sliding along it much faster than the impact can stop: the contact
    slides throughout, so the tangential impulse is the Coulomb limit, |J_t| = mu |J_n| with b2MixFriction's
    mu = sqrt(mu_a mu_b), for both bodies (equal and opposite impulses)."""
    env = _two_box_env(0.0)
    env["objects"][0]["links"][0]["contact_friction"] = mu_a
    env["objects"][1]["links"][0]["contact_friction"] = mu_b
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[1.0, 2.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.3, 0.05, 0.0], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    w.step(40)
    _, qd = w.get_state()
    mu = np.sqrt(mu_a * mu_b)
    jn_a, jt_a = 0.5 * (qd[0] - 1.0), 0.5 * (qd[1] - 2.0)
    jn_b, jt_b = 0.5 * qd[3], 0.5 * qd[4]
    assert jn_a < -0.05 and jn_b > 0.05                                  # a real impact
    assert jt_a / jn_a == pytest.approx(mu, rel=1e-4) and jt_b / jn_b == pytest.approx(mu, rel=1e-4)
    assert jn_a + jn_b == pytest.approx(0.0, abs=1e-6) and jt_a + jt_b == pytest.approx(0.0, abs=1e-6)


def test_fat_aabb_follows_the_move_proxy_rule(oracle):
    """b2Body::SynchronizeFixtures + b2DynamicTree::MoveProxy: a proxy keeps its fat box while the swept tight box (polygon
    +- b2_polygonRadius at the old and the new transform) stays inside; otherwise the new fat box is the swept box
    +- b2_aabbExtension (0.1) stretched by b2_aabbMultiplier (2) x displacement on the leading side.  A 2 x 2 box (Box2D
    units) moving at 0.3 units per step: refits at step 1 to [-1.11, 1.01 + 0.3 + 0.1 + 0.6] and again at step 4."""
    env = _two_box_env(0.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[3.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.0, 3.0, 0.0], velocity=[0.0, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    fat = w.fat_aabbs()
    assert np.allclose(fat[1], [-1.11, -1.11, 1.11, 1.11], atol=1e-6)          # creation: tight box +- 0.1
    assert np.allclose(fat[2], [-1.11, 28.89, 1.11, 31.11], atol=1e-5)
    expect = {1: [-1.11, 2.01], 2: [-1.11, 2.01], 3: [-1.11, 2.01], 4: [-0.21, 2.91], 5: [-0.21, 2.91]}
    for step in range(1, 6):
        w.step(1)
        f = w.fat_aabbs()[1]
        assert np.allclose(f[[0, 2]], expect[step], atol=1e-5), step
        assert np.allclose(f[[1, 3]], [-1.11, 1.11], atol=1e-6)
    assert np.array_equal(w.fat_aabbs()[2], fat[2])                            # the resting box never refits


def test_sleep_thresholds(oracle):
    """b2Island::Solve's sleep rule: a body slower than b2_linearSleepTolerance (0.01 Box2D units/s) and
    b2_angularSleepTolerance accumulates sleep time; the island is put to sleep (velocity zeroed) once the minimum reaches
    b2_timeToSleep = 0.5 s -- in float arithmetic 50 x 0.01 is 0.4999998, so that is step 51.  A body at twice the tolerance
    never sleeps; stepping with allow_sleeping = false keeps the slow one awake too."""
    env = _two_box_env(0.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[0.0005, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.0, 3.0, 0.0], velocity=[0.002, 0.0, 0.0])
    w = oracle.OracleWorld(env)
    w.step(50)
    b = w.bodies()
    assert b[1, 14] == 1.0 and b[1, 10] == pytest.approx(0.005, rel=1e-6) and b[1, 13] == pytest.approx(0.5, abs=1e-6) and b[1, 13] < 0.5
    w.step(1)
    b = w.bodies()
    assert b[1, 14] == 0.0 and b[1, 10] == 0.0 and b[1, 13] == 0.0            # asleep, velocity and sleep time reset
    assert b[2, 14] == 1.0 and b[2, 10] == pytest.approx(0.02, rel=1e-6)      # the faster box stays awake
    w2 = oracle.OracleWorld(env)
    w2.step(80, True, False)
    b = w2.bodies()
    assert b[1, 14] == 1.0 and b[1, 10] == pytest.approx(0.005, rel=1e-6)


def test_revolute_limit_is_an_inelastic_stop(oracle):
    """b2RevoluteJoint with enableLimit: the limit only acts once the angle has reached it (so the arm overshoots by at most
    one step of travel), then the velocity constraint removes the whole relative speed (no restitution on joint limits) and
    the position constraint tolerates b2_angularSlop (2 degrees) beyond the limit: the arm parks within
    [upper, upper + angularSlop]."""
    def link(name, poly, mass):
        return dict(name=name, geometry=[poly], mass=mass, ground_friction=0.0, ground_torque_integral=0.0, contact_friction=0.0,
                    restitution=0.0, balls=[])
    robot = dict(name="r", static=False,
                 links=[link("base", [-0.5, -0.5, 0.5, -0.5, 0.5, 0.5, -0.5, 0.5], 3000.0),
                        link("arm", [-0.05, 0.0, 0.05, 0.0, 0.05, 0.4, -0.05, 0.4], 0.3)],
                 joints=[dict(name="hinge", link_a="base", link_b="arm", axis=[0.0, 0.6], axis_orientation=0.0, position_limits=[-0.5, 0.5],
                              velocity_limits=[-100.0, 100.0], acceleration_limits=[-1000.0, 1000.0], joint_type="revolute",
                              actuated=False)],
                 base_actuation=dict(translational_velocity_limits=[-2.0, 2.0], translational_acceleration_limits=[-2.0, 2.0],
                                     rotational_velocity_limits=[-2.0, 2.0], rotational_acceleration_limits=[-2.0, 2.0],
                                     use_center_of_mass=True))
    env = dict(scale=10.0, world_bounds=[-5, -5, 5, 5], robots=[robot], objects=[],
               states={"r": dict(configuration=[0.0, 0.0, 0.0, 0.0], velocity=[0.0, 0.0, 0.0, 2.0])})
    w = oracle.OracleWorld(env)
    w.step(24)
    q, qd = w.get_state()
    assert q[3] == pytest.approx(24 * 0.02, rel=2e-3) and qd[3] == pytest.approx(2.0, rel=5e-3)   # free swing below the limit
    w.step(16)
    q, qd = w.get_state()
    slop = 2.0 / 180.0 * np.pi
    assert 0.5 <= q[3] <= 0.5 + min(slop, 0.02 + 1e-4) and abs(qd[3]) < 1e-6
    w.step(50)
    q2, _ = w.get_state()
    assert q2[3] == pytest.approx(q[3], abs=1e-5)                                                  # and stays there


def test_toi_lands_the_box_at_the_wall(oracle):
    """Continuous collision against a static body: a box flying 1.5 Box2D units per step at a wall 0.5 units away is
    advanced to the time of impact (core shapes b2_linearSlop = 0.005 apart, tolerance a quarter of it), b2Island::SolveTOI
    pushes the two manifold points out to -1.5 linearSlop of skinned separation (b2_toiBaugarte 0.75: +0.0075, then
    +0.0019 for the second point) and removes the approach speed; the rest of the step is spent at rest.  It never
    penetrates, lands between 0.0125 and 0.016 units from the wall and creeps to the resting gap
    2 polygonRadius - linearSlop = 0.015 afterwards."""
    env = _two_box_env(0.0)
    env["objects"][1]["static"] = True
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[15.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[1.0, 0.02, 0.0], velocity=[])
    w = oracle.OracleWorld(env)

    def gap():
        b = w.bodies()
        return float((b[2, 4] - 1.0) - (b[1, 4] + 1.0))      # between the facing sides, Box2D units (half width 1)

    for step in range(1, 6):
        w.step(1)
        assert w.stats()["toi_events"] == 0 and gap() == pytest.approx(8.0 - 1.5 * step, abs=1e-4)
    w.step(1)
    _, qd = w.get_state()
    # the approach speed is gone; the frictionless two-point impact leaves a sideways drift of under a millimetre per second
    assert w.stats()["toi_events"] == 1 and abs(qd[0]) < 1e-5 and abs(qd[1]) < 5e-3
    assert 0.005 + 0.0075 - 0.00125 <= gap() <= 0.016
    landed = gap()
    w.step(200)
    assert landed <= gap() <= 0.015 + 1e-5 and gap() == pytest.approx(0.015, abs=2e-4)


def test_forward_kinematics_known_answers(oracle):
    """setDOFPositions -> link poses (Box2DJoint::resetPosition :992-1019, Box2DLink::setPose :405-438): a child link's origin
    sits at its joint's axis point in the parent frame and is turned by the joint angle plus axis_orientation; the robot's
    pose DOFs are the base link's centre of mass (its origin was moved there, :653-671)."""
    env = load_env_dict(ENV_A)
    w = oracle.OracleWorld(env)
    q, qd = w.get_state()
    q[:7] = [0.4, -0.3, 0.7, 0.5, -0.4, -0.6, 0.3]
    w.set_state(q, np.zeros_like(qd))
    ls = w.link_states().astype(np.float64)
    bm = w.body_model().astype(np.float64)
    joints = {j["name"]: j for j in env["robots"][0]["joints"]}

    def rot(a):
        return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])

    base_angle = 0.7
    base_origin = np.array([0.4, -0.3]) - rot(base_angle) @ (bm[1, 4:6] / 10.0)      # body origin = COM - R * local COM
    assert ls[1, 2] == pytest.approx(base_angle, abs=1e-6) and np.allclose(ls[1, :2], [0.4, -0.3], atol=1e-6)
    chain = [("finger_l_joint_1", 2, 1, 0.5), ("finger_l_joint_2", 3, 2, -0.4), ("finger_r_joint_1", 4, 1, -0.6),
             ("finger_r_joint_2", 5, 4, 0.3)]
    origin, angle = {1: base_origin}, {1: base_angle}
    for name, child, parent, value in chain:
        jd = joints[name]
        origin[child] = origin[parent] + rot(angle[parent]) @ np.array(jd["axis"], np.float64)
        angle[child] = angle[parent] + value + jd["axis_orientation"]
        assert np.allclose(ls[child, :2], origin[child], atol=2e-6), name
        assert ls[child, 2] == pytest.approx(angle[child], abs=2e-6), name
    q1, _ = w.get_state()
    assert np.allclose(q1[:7], q[:7], atol=2e-6)                                     # and the DOFs read back


def test_velocity_kinematics_known_answers(oracle):
    """setDOFVelocities -> body velocities (Box2DObject::updateBodyVelocities :2018-2044, Box2DLink::updateBodyVelocities
    :759-800): every link moves with the base (translation plus rotation about the base link's centre of mass) and turns
    about each joint above it: v_com = v_base + w_base x (com - com_base) + sum_i qd_i x (com - anchor_i), w = w_base + sum_i qd_i."""
    env = load_env_dict(ENV_A)
    w = oracle.OracleWorld(env)
    q, qd = w.get_state()
    q[:7] = [0.1, 0.2, -0.5, 0.3, 0.2, -0.2, -0.3]
    qd[:7] = [0.3, -0.1, 0.8, 1.5, -2.0, 0.7, 1.1]
    w.set_state(q, qd)
    ls = w.link_states().astype(np.float64)

    def cross(omega, r):
        return omega * np.array([-r[1], r[0]])

    v_base, w_base, com_base = np.array([0.3, -0.1]), 0.8, ls[1, 6:8]
    assert np.allclose(ls[1, 3:6], [0.3, -0.1, 0.8], atol=1e-6)
    # the joint anchors are the child links' origins (pose), the chains are base -> l1 -> l2 and base -> r1 -> r2
    chains = {2: [(2, 1.5)], 3: [(2, 1.5), (3, -2.0)], 4: [(4, 0.7)], 5: [(4, 0.7), (5, 1.1)]}
    for link, joints in chains.items():
        com = ls[link, 6:8]
        v = v_base + cross(w_base, com - com_base)
        omega = w_base
        for anchor_link, rate in joints:
            v = v + cross(rate, com - ls[anchor_link, :2])
            omega += rate
        assert np.allclose(ls[link, 3:5], v, atol=5e-6), link
        assert ls[link, 5] == pytest.approx(omega, abs=1e-6), link
    _, qd1 = w.get_state()
    assert np.allclose(qd1[:7], qd[:7], atol=5e-6)


def test_check_collision_known_answers(oracle):
    """Box2DCollisionChecker (Box2DWorld.cpp:2832-2911): polygons touch when their cores are closer than the two polygon
    radii (0.02 Box2D units = 2 mm); the reported point is b2WorldManifold::points[0] -- midway between the skinned surfaces,
    at the first incident vertex -- and the reference multiplies point AND unit normal by 1 / scale (:2857-2862)."""
    env = _two_box_env(0.0)
    env["states"]["a"] = dict(configuration=[0.0, 0.0, 0.0], velocity=[0.0, 0.0, 0.0])
    env["states"]["b"] = dict(configuration=[0.201, 0.05, 0.0], velocity=[0.0, 0.0, 0.0])      # faces 1 mm apart
    io, fo = oracle.OracleWorld(env).check_collision()
    assert io.tolist() == [[1, 2, 1, 2]]                                                      # bodies a, b; their fixtures
    assert np.allclose(fo[0, :3], [0.1005, -0.05, 0.0], atol=1e-6)                            # mid-gap, B's lower corner
    assert np.allclose(fo[0, 3:], [0.1, 0.0, 0.0], atol=1e-7)                                 # unit normal / scale
    env["states"]["b"] = dict(configuration=[0.2021, 0.05, 0.0], velocity=[0.0, 0.0, 0.0])     # 2.1 mm: beyond the radii
    io, fo = oracle.OracleWorld(env).check_collision()
    assert len(io) == 0


def test_recorded_contacts_known_answers(oracle):
    """stepPhysics(std::vector<Contact>&, steps) (Box2DWorld.cpp:3290-3310, BeginContact :2792-2810): one record per
    BeginContact.  Box a (1 m/s) meets resting box b whose near face is at x = 0.4 m: exactly one event, at that face,
    with the unit normal from a to b divided by the scale; after the elastic bounce they separate and nothing else begins."""
    w = oracle.OracleWorld(_two_box_env(1.0))
    io, fo = w.step_record(60)
    assert io.tolist() == [[1, 2, 1, 2]]
    assert fo[0, 0] == pytest.approx(0.4, abs=2e-3) and abs(fo[0, 1]) == pytest.approx(0.1, abs=1e-6) and fo[0, 2] == 0.0
    assert np.allclose(fo[0, 3:], [0.1, 0.0, 0.0], atol=1e-7)
    io, fo = w.step_record(20)
    assert len(io) == 0


# ---- b2PrismaticJoint (row 8f-4): known answers from the constraint definitions --------------------------------------------------
def _rail(world):
    """(translation, lateral offset, relative angle, limit state, motor impulse) of the drawer's prismatic joint from body states."""
    b = world.bodies()
    cab, sl = b[-2], b[-1]
    ji, jf = world.joints()
    ax = np.array([cab[3], cab[2]])                 # cabinet x axis in the world (cos, sin)
    perp = np.array([-cab[2], cab[3]])
    anchor = cab[0:2] + 3.5 * ax                    # localAnchorA = (3.5, 0) on the cabinet, Box2D units
    d = sl[0:2] - anchor                            # localAnchorB = (0, 0) on the slider
    return float(d @ ax), float(d @ perp), float(sl[6] - cab[6]), int(ji[-1][3]), float(jf[-1][3])


def test_prismatic_joint_known_answers(oracle):
    from parity_utils import env_prismatic
    slop, ang_slop = 0.005, 2.0 / 180.0 * np.pi
    # (a) free rail: the joint removes the lateral and the angular degree of freedom while the robot rams the drawer
    w = oracle.OracleWorld(oracle.OracleEnv(env_prismatic()))
    assert abs(w.get_state()[0][-1] - 0.5) < 1e-6       # created 0.5 units inside the travel; getPosition is in Box2D units
    w.set_target_velocity(0, np.array([1.0, 0, 0, 0, 0, 0, 0], np.float32))
    moved = 0.0
    for _ in range(12):
        w.step(10)
        t, lat, ang, state, _ = _rail(w)
        assert abs(lat) <= 3 * slop and abs(ang) <= 2 * ang_slop, (lat, ang)
        assert -1.0 - 2 * slop <= t <= 2.0 + 2 * slop   # limits [-0.1, 0.2] m are scaled by 10
        assert abs(t - w.get_state()[0][-1]) < 1e-4
        moved = max(moved, abs(t - 0.5))
    assert moved > 0.2, "the robot never moved the slider"
    # (b) created below its lower limit: the position solver pushes the slider to lower - linearSlop and holds it there
    w = oracle.OracleWorld(oracle.OracleEnv(env_prismatic(limits=(0.1, 0.2))))
    w.step(10)
    t, lat, ang, state, _ = _rail(w)
    assert state == 1 and abs(t - (1.0 - slop)) < 1e-3, (t, state)
    # (c) equal limits: b2PrismaticJoint pins the TRANSLATION to zero (C = translation), not to the limit value
    w = oracle.OracleWorld(oracle.OracleEnv(env_prismatic(limits=(0.05, 0.0501))))
    w.step(30)
    t, lat, ang, state, _ = _rail(w)
    assert state == 3 and abs(t) < 2 * slop, (t, state)
    # (d) actuated: enableMotor with motorSpeed 0 and maxMotorForce = min |acceleration limits| = 0.6 -> a brake whose impulse
    # saturates at dt * maxMotorForce while the slider is being pushed
    w = oracle.OracleWorld(oracle.OracleEnv(env_prismatic(actuated=True)))
    w.set_target_velocity(0, np.array([1.0, 0, 0, 0, 0, 0, 0], np.float32))
    peak = 0.0
    for _ in range(90):
        w.step(1)
        peak = max(peak, abs(_rail(w)[4]))
    assert abs(peak - 0.01 * 0.6) < 1e-6, peak
    # (e) what the reference cannot do, it refuses: placing the drawer, a `states` entry for it, a robot with a prismatic joint
    w = oracle.OracleWorld(oracle.OracleEnv(env_prismatic()))
    q, qd = w.get_state()
    with pytest.raises(Exception):
        w.set_state(q, qd)
    bad = env_prismatic()
    bad["states"]["drawer"] = dict(configuration=[1.0, 1.0, 0.0, 0.0], velocity=[0.0] * 4)
    with pytest.raises(Exception):
        oracle.OracleWorld(oracle.OracleEnv(bad))
