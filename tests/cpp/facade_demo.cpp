// Exercises include/b2g_world.hpp the way a planner built on sim_env::Box2DWorld would (C++11, no CUDA headers needed).
//   facade_demo <env.yaml> host   -> model introspection works, creating the batch fails loudly without a GPU
//   facade_demo <env.yaml> gpu    -> steps 64 worlds and prints world 0's state as hex floats for the Python-side check
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "b2g_world.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string path = argv[1];
    const bool gpu = std::strcmp(argv[2], "gpu") == 0;
    sim_env::b2g::BatchBox2DWorld batch(64, 0);
    try {
        batch.loadWorld(path);
    } catch (const std::runtime_error& e) {
        if (gpu) {
            std::printf("FAIL %s\n", e.what());
            return 1;
        }
        // no CUDA device: the model stage must still have succeeded and the failure must name CUDA
        bool named = std::string(e.what()).find("cuda") != std::string::npos || std::string(e.what()).find("CUDA") != std::string::npos;
        // name lookups are model-only: Box2DObject::getLink / getJoint / getBaseLink
        int lookup = 0;
        try {
            lookup = batch.getLink("obj1", "test_object_link") == 6 && batch.getBaseLink("my_robot") == 1 &&
                     batch.getLinkName(3) == "test_robot_finger_l_2" && batch.getLinks("obj2").size() == 2 &&
                     batch.getJointInfo(batch.getJoint("my_robot", "finger_r_joint_2")).dof_index == 6;
            try {
                batch.getLink("obj1", "no_such_link");
                lookup = 0;
            } catch (const std::runtime_error&) {
            }
        } catch (const std::exception&) {
            lookup = 0;
        }
        std::printf("HOST nq=%d bodies=%d objects=%zu first=%s cuda_named=%d lookup=%d\n", batch.nq(), batch.numBodies(),
                    batch.getObjects().size(), batch.getObjects().empty() ? "-" : batch.getObjects()[0].name.c_str(), (int)named, lookup);
        return named && lookup ? 0 : 1;
    }
    if (!gpu) {
        std::printf("HOST nq=%d bodies=%d objects=%zu first=%s cuda_named=-1\n", batch.nq(), batch.numBodies(), batch.getObjects().size(),
                    batch.getObjects()[0].name.c_str());
        return 0;
    }
    const int n = (int)batch.numWorlds(), nq = batch.nq();
    std::vector<float> q, qd;
    batch.getWorldState(q, qd);
    for (int w = 0; w < n; ++w) q[(size_t)w * nq + 7] = 0.2f + 0.01f * w;  // obj1 x: slides it towards the robot's fingers
    batch.setWorldState(q, qd, true);
    const int nd = batch.robotDofs("my_robot");
    std::vector<float> u((size_t)n * nd, 0.0f);
    for (int w = 0; w < n; ++w) { u[(size_t)w * nd] = 0.5f; u[(size_t)w * nd + 2] = 0.3f; }
    batch.setTargetVelocity("my_robot", u.data());
    batch.stepPhysics(10);
    batch.saveState();
    std::vector<sim_env::b2g::Contact> contacts;
    std::vector<int64_t> offsets;
    batch.stepPhysics(contacts, offsets, 20);
    size_t recorded = contacts.size();
    bool restored = batch.restoreState();
    batch.stepPhysics(20);
    batch.getWorldState(q, qd);
    std::vector<bool> hit = batch.checkCollision();
    int hits = 0;
    for (int w = 0; w < n; ++w) hits += hit[w] ? 1 : 0;
    std::printf("GPU recorded=%zu restored=%d hits=%d q0=", recorded, (int)restored, hits);
    for (int i = 0; i < nq; ++i) std::printf("%a ", q[i]);
    std::printf("\n");
    bool threw = false;
    try {
        std::vector<float> bad(3);
        batch.setWorldState(bad, bad);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    // ---- single-world views: the one-world signatures of Box2DWorld / Box2DObject on world 3
    using namespace sim_env::b2g;
    int bad = 0;
    World w3 = batch.world(3);
    Object robot = w3.getRobot("my_robot"), obj1 = w3.getObject("obj1"), obj2 = w3.getObject("obj2");
    bad += !batch.isRobot("my_robot") || batch.isRobot("obj1") || batch.isRobot("nobody") || !batch.supportsPhysics();
    bad += std::fabs(batch.getGravity() - 98.1f) > 1e-4f;
    bad += robot.getNumDOFs() != 7 || obj1.getNumDOFs() != 3 || obj2.getNumDOFs() != 1 || !obj2.isStatic();
    bad += robot.getType() != EntityType::Robot || obj1.getType() != EntityType::Object;
    bad += !robot.getDOFInformation(2).cyclic || robot.getDOFInformation(3).cyclic;
    std::vector<float> p = obj1.getDOFPositions();  // active DOFs = all DOFs
    for (int i = 0; i < 3; ++i) bad += p[i] != q[(size_t)3 * nq + 7 + i];
    std::vector<float> rv = robot.getDOFVelocities(std::vector<int>{0, 2, 5});
    bad += rv[0] != qd[(size_t)3 * nq + 0] || rv[1] != qd[(size_t)3 * nq + 2] || rv[2] != qd[(size_t)3 * nq + 5];
    obj1.setDOFPositions(std::vector<float>{0.55f}, std::vector<int>{0});  // x only, world 3 only
    std::vector<float> q2, qd2;
    batch.getWorldState(q2, qd2);
    bad += std::fabs(q2[(size_t)3 * nq + 7] - 0.55f) > 1e-5f || std::fabs(q2[(size_t)3 * nq + 8] - q[(size_t)3 * nq + 8]) > 1e-5f;
    for (int i = 0; i < nq; ++i) bad += q2[(size_t)4 * nq + i] != q[(size_t)4 * nq + i] || q2[(size_t)2 * nq + i] != q[(size_t)2 * nq + i];
    WorldState ws = w3.getWorldState();
    bad += ws.size() != 3 || std::fabs(ws["obj1"].dof_positions[0] - 0.55f) > 1e-5f;
    ws["obj1"].dof_positions[1] = -0.25f;
    ws["obj1"].pose[1] = -0.25f;
    bad += !w3.setWorldState(ws);
    bad += std::fabs(obj1.getDOFPositions(std::vector<int>{1})[0] + 0.25f) > 1e-5f;
    WorldState unknown;
    unknown["nobody"] = ws["obj1"];
    bad += w3.setWorldState(unknown);  // unknown names: false, like the reference (:3505-3510)
    // collision overloads agree with the world's contact list
    int viewHits = 0;
    for (int w = 0; w < n; ++w) {
        World vw = batch.world(w);
        std::vector<Contact> all, ro;
        bool any = vw.checkCollision(all);
        Object r = vw.getRobot("my_robot"), o1 = vw.getObject("obj1");
        bool pair = vw.checkCollision(r, o1, ro);
        bool expect = false;
        size_t cnt = 0;
        const int rb0 = batch.getObjects()[0].first_body, rn = batch.getObjects()[0].n_links;
        const int ob0 = batch.getObjects()[1].first_body, on = batch.getObjects()[1].n_links;
        for (size_t k = 0; k < all.size(); ++k) {
            bool aR = all[k].body_a >= rb0 && all[k].body_a < rb0 + rn, bR = all[k].body_b >= rb0 && all[k].body_b < rb0 + rn;
            bool aO = all[k].body_a >= ob0 && all[k].body_a < ob0 + on, bO = all[k].body_b >= ob0 && all[k].body_b < ob0 + on;
            if ((aR && bO) || (aO && bR)) { expect = true; ++cnt; }
        }
        bad += pair != expect || ro.size() != cnt || (any != !all.empty());
        for (size_t k = 0; k < ro.size(); ++k) bad += !(ro[k].body_a >= rb0 && ro[k].body_a < rb0 + rn);  // link_a is the queried one
        bad += r.checkCollision(o1) != pair || o1.checkCollision(r) != pair;
        std::vector<Object> collidors;
        bool rc = vw.checkCollision(r, collidors);
        bad += rc != r.checkCollision();
        viewHits += pair ? 1 : 0;
    }
    // clone(): same parameters and state, independent afterwards
    {
        BatchBox2DWorld* twin = batch.clone();
        std::vector<float> tq, tqd, bq, bqd;
        twin->getWorldState(tq, tqd);
        batch.getWorldState(bq, bqd);
        for (size_t i = 0; i < tq.size(); ++i) bad += std::fabs(tq[i] - bq[i]) > 1e-5f;
        bad += twin->getPhysicsTimeStep() != batch.getPhysicsTimeStep();
        twin->stepPhysics(3);
        batch.getWorldState(tq, tqd);
        for (size_t i = 0; i < tq.size(); ++i) bad += tq[i] != bq[i];  // stepping the clone leaves the original alone
        delete twin;
    }
    // Box2DLink's dynamics parameters on the shared model: getters follow, the clone keeps the loaded values
    {
        const int body = batch.getLink("obj1", "test_object_link");  // obj1's only link
        bad += body != 6 || batch.getBaseLink("obj1") != body || batch.getLinkName(body) != "test_object_link";
        bad += batch.getLinks("my_robot").size() != 5 || batch.getBaseLink("my_robot") != batch.getLink("my_robot", "test_robot_base");
        const b2g_joint_info ji = batch.getJointInfo(batch.getJoint("my_robot", "finger_r_joint_2"));
        bad += ji.dof_index != 6 || ji.body_b != batch.getLink("my_robot", "test_robot_finger_r_2") || !ji.actuated;
        const float m0 = batch.getMass(body), i0 = batch.getInertia(body);
        batch.setMass(body, 2.0f * m0);
        batch.setGroundFrictionCoefficient(body, 0.5f);
        batch.setGroundFrictionTorqueIntegral(body, 0.04f);
        batch.setContactFriction(body, 0.7f);
        bad += std::fabs(batch.getMass(body) - 2.0f * m0) > 1e-6f || std::fabs(batch.getInertia(body) - 2.0f * i0) > 1e-6f;
        bad += std::fabs(batch.getGroundFrictionLimitForce(body) - 0.5f * 2.0f * m0 * 9.81f) > 1e-5f;
        bad += std::fabs(batch.getGroundFrictionLimitTorque(body) - 0.5f * 0.04f) > 1e-7f || batch.getContactFriction(body) != 0.7f;
        std::vector<float> links;
        batch.getLinkStates(links);
        float pose[3];
        batch.world(5).getObject("obj1").getPose(pose);
        const float* row = &links[((size_t)5 * batch.numBodies() + body) * 8];
        bad += links.size() != (size_t)n * batch.numBodies() * 8 || row[0] != pose[0] || row[1] != pose[1] || row[2] != pose[2];
        BatchBox2DWorld* twin = batch.clone();
        bad += twin->getMass(body) != m0;
        delete twin;
        // per-world dynamics at tile granularity: the second tile of 32 worlds gets its own mass for obj1
        if (n >= 64) {
            const float mAll = batch.getMass(body);
            batch.setLinkPropertyOfWorlds(32, 32, body, B2G_LINK_MASS, 3.0f * m0);
            bad += batch.getLinkPropertyOfWorld(40, body, B2G_LINK_MASS) != 3.0f * m0;
            bad += batch.getLinkPropertyOfWorld(5, body, B2G_LINK_MASS) != mAll || batch.getLinkPropertyOfWorld(31, body, B2G_LINK_MASS) != mAll;
            bool insideTileThrew = false;
            try {
                batch.setLinkPropertyOfWorlds(33, 4, body, B2G_LINK_MASS, 1.0f);
            } catch (const std::exception&) {
                insideTileThrew = true;
            }
            bad += !insideTileThrew;
        }
        batch.stepPhysics(2);
        bool groundThrew = false;
        try {
            batch.setMass(0, 1.0f);
        } catch (const std::exception&) {
            groundThrew = true;
        }
        bad += !groundThrew;
    }
    // ---- Link / Joint views (Box2DLink / Box2DJoint) and the logger shim
    {
        Link base = robot.getBaseLink(), tip = robot.getLink("test_robot_finger_l_2");
        Joint j = robot.getJoint("finger_l_joint_2"), parent;
        bad += base.getName() != "test_robot_base" || tip.body() != 3 || robot.getLinks().size() != 5 || robot.getJoints().size() != 4;
        bad += !tip.getParentJoint(parent) || parent.index() != j.index() || base.getParentJoint(parent) || base.getChildJoints().size() != 2;
        bad += j.getJointType() != Joint::JointType::Revolute || j.getDOFIndex() != 4 || j.getJointIndex() != 1;
        bad += j.getChildLink().body() != tip.body() || j.getParentLink().getName() != "test_robot_finger_l_1";
        float lim[2];
        j.getPositionLimits(lim);
        bad += std::fabs(lim[0] + 1.05f) > 1e-6f || std::fabs(lim[1] - 1.05f) > 1e-6f;
        // a joint's position / velocity are one DOF of its object, in this world only
        const std::vector<float> before = robot.getDOFPositions();
        j.setPosition(0.4f);
        j.setVelocity(-0.2f);
        const std::vector<float> after = robot.getDOFPositions(), w4q = batch.world(4).getRobot("my_robot").getDOFPositions();
        bad += std::fabs(j.getPosition() - 0.4f) > 1e-5f || std::fabs(j.getVelocity() + 0.2f) > 1e-5f || std::fabs(after[4] - 0.4f) > 1e-5f;
        bad += std::fabs(after[3] - before[3]) > 1e-6f || std::fabs(w4q[4] - 0.4f) < 1e-3f;
        j.setPosition(before[4]);
        // link pose: the base link's pose is the robot's first three DOFs; mass as in the description
        float pose[3], com[2], vel[3];
        base.getPose(pose);
        base.getCenterOfMass(com);
        base.getVelocityVector(vel);
        bad += std::fabs(pose[0] - after[0]) > 1e-6f || std::fabs(pose[1] - after[1]) > 1e-6f || std::fabs(pose[2] - after[2]) > 1e-6f;
        bad += std::fabs(base.getMass() - 10.1f) > 1e-5f || std::fabs(tip.getMass() - 0.1f) > 1e-6f || tip.getContactFriction() != 1.5f;
        bad += (base.checkCollision() || tip.checkCollision()) != robot.checkCollision();
        static int logged;
        logged = 0;
        w3.getLogger().setSink([](Logger::LogLevel lvl, const std::string&, const std::string&) { logged += lvl == Logger::LogLevel::Warn; });
        w3.getLogger().logDebug("below the level", "[facade_demo]");
        w3.getLogger().logWarn("a warning", "[facade_demo]");
        bad += logged != 1;
    }
    bool viewThrew = false;
    try {
        w3.stepPhysics(1);
    } catch (const std::logic_error&) {
        viewThrew = true;
    }
    std::printf("VIEW bad=%d hits=%d threw=%d\n", bad, viewHits, (int)viewThrew);
    return threw && viewThrew && bad == 0 ? 0 : 1;
}
