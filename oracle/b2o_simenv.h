// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// CPU restatement of the sim_env wrapper semantics of the reference on top of the b2o engine:
// model construction (/root/reference/src/sim_env/Box2DWorld.cpp:44-148, 872-948, 1292-1363,
// 2046-2067, 3644-3721), DOF <-> body kinematics (:405-438, 759-794, 992-1028, 1507-1715,
// 2018-2041), the velocity controller (/root/reference/src/sim_env/Box2DController.cpp:42-117,
// 191-215), effort application (:1198-1222, 2284-2379), stepping (:3266-3333) and the collision
// checker (:2743-2911).
#pragma once
#include <stdexcept>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "b2o_dynamics.h"

namespace b2o {

struct LinkDesc {
    std::string name;
    std::vector<std::vector<float>> polygons;  // flat x,y lists (unscaled)
    float mass = 0.0f;
    float ground_friction = 0.0f;
    float ground_torque_integral = 0.0f;
    float contact_friction = 0.0f;
    float restitution = 0.0f;
    std::vector<std::vector<float>> balls;  // [x, y, z, radius] in the link frame (Box2DIOUtils.h:21, 139-142)
};
struct JointDesc {
    std::string name, link_a, link_b;
    float axis[2] = {0, 0};
    float position_limits[2] = {0, 0};
    float velocity_limits[2] = {0, 0};
    float acceleration_limits[2] = {0, 0};
    std::string joint_type = "revolute";
    bool actuated = false;
    float axis_orientation = 0.0f;
};
struct ObjectDesc {
    std::string name;
    std::string base_link;
    std::vector<LinkDesc> links;
    std::vector<JointDesc> joints;
    bool is_static = false;
    bool is_robot = false;
    float translational_velocity_limits[2] = {0, 0};
    float translational_acceleration_limits[2] = {0, 0};
    float rotational_velocity_limits[2] = {0, 0};
    float rotational_acceleration_limits[2] = {0, 0};
    bool use_center_of_mass = true;
};
struct StateDesc {
    std::vector<float> configuration, velocity;
};
struct EnvDesc {
    float scale = 1.0f;
    float world_bounds[4] = {0, 0, 0, 0};
    std::vector<ObjectDesc> robots, objects;
    std::map<std::string, StateDesc> states;
};

struct OWorld;
struct OObject;
struct OJoint;

struct OLink {
    std::string name;
    OWorld* world = nullptr;
    OObject* object = nullptr;
    Body* body = nullptr;
    Joint* frictionJoint = nullptr;
    Vec2 localOriginOffset;
    float localAabbMin[2], localAabbMax[2];
    float contactFriction = 0, groundFriction = 0, groundTorqueIntegral = 0;
    std::vector<OJoint*> childJoints;
    OJoint* parentJoint = nullptr;
    bool enabled = true;
    std::vector<std::vector<float>> balls;  // Box2DLink::_balls (:54-57): [x, y, z, radius]

    void getPose(float pose[3]) const;
    // Box2DLink::updateBallApproximation (:733-747): appends [x, y, z, radius] with centre = getTransform() * centre
    void ballApproximation(std::vector<float>& out) const;
    void setPose(const float pose[3], bool updateChildren = true, bool jointOverride = false);
    void getVelocityVector(float v[3]) const;
    float getMass() const { return body->GetMass(); }
    float getInertia() const;
    float getCOMInertia() const;
    void getCenterOfMass(float com[2]) const;
    void setOriginToCenterOfMass();
    void updateBodyVelocities(const std::vector<float>& allDofVel, std::vector<std::pair<Vec2, unsigned>>& parentJoints,
                              unsigned indexOffset);
    void setEnabled(bool e);
    // dynamics parameters (Box2DWorld.cpp:478-547)
    void setMass(float mass);
    void setGroundFrictionCoefficient(float mu) { groundFriction = mu; updateFrictionJoint(); }
    void setGroundFrictionTorqueIntegral(float v) { groundTorqueIntegral = v; updateFrictionJoint(); }
    void setContactFriction(float mu);
    void updateFrictionJoint();
    float getGroundFrictionLimitTorque() const { return groundFriction * groundTorqueIntegral; }  // :502-505
    float getGroundFrictionLimitForce() const { return groundFriction * getMass() * 9.81f; }      // :507-510
};

struct OJoint {
    std::string name;
    OWorld* world = nullptr;
    OObject* object = nullptr;
    OLink* linkA = nullptr;
    OLink* linkB = nullptr;
    Joint* joint = nullptr;
    unsigned jointIndex = 0, dofIndex = 0;
    float positionLimits[2], velocityLimits[2], accelerationLimits[2];

    bool prismatic() const { return joint->kind == kPrismaticJoint; }
    // Box2DJoint::getPosition / getVelocity (:965-981, :1030-1046): a prismatic joint reports b2PrismaticJoint's translation /
    // speed as they are, i.e. in Box2D units (the reference does not scale them back)
    float getPosition() const { return prismatic() ? joint->GetJointTranslation() : joint->GetJointAngle(); }
    float getVelocity() const { return prismatic() ? joint->GetPrismaticJointSpeed() : joint->GetJointSpeed(); }
    void resetPosition(float value, bool childJointOverride);
    void setControlTorque(float value);
    void getAxisPosition(float axis[2]) const;
    Vec2 getGlobalAxisPosition() const {  // :1274-1287
        if (prismatic()) throw std::logic_error("getLocalAxisPosition for prismatic joints is not implemented yet");
        return joint->GetAnchorA();
    }
};

struct DofInfo {
    float positionLimits[2], velocityLimits[2], accelerationLimits[2];
};

struct OObject {
    std::string name;
    OWorld* world = nullptr;
    bool isStatic = false;
    bool isRobot = false;
    float mass = 0.0f;
    float localAabb[4] = {0, 0, 0, 0};     // Box2DObject::_local_bounding_box (:1304-1312): min x, min y, max x, max y
    std::map<std::string, OLink*> links;   // name order, as the reference's std::map
    std::vector<OLink*> linksCreation;     // YAML order
    std::vector<OJoint*> sortedJoints;     // YAML order
    OLink* baseLink = nullptr;
    unsigned numDofs = 0;
    std::vector<int> activeDofs;
    DofInfo baseDof[3];
    // velocity controller state (Box2DRobotVelocityController)
    std::vector<float> targetVelocity;
    bool targetAvailable = false;
    // a ControlCallback that returns fixed efforts (Box2DRobot::setController :2312-2317): control() hands them to
    // commandEfforts every step
    bool effortController = false;
    std::vector<float> controlEfforts;
    // position controller on top of the velocity controller (sim_env::RobotPositionController / SE2RobotPositionController,
    // instantiated at Box2DWorldViewer.cpp:1208-1216; the classes live in the absent sim_env package).  ASSUMED law, the same
    // on the CUDA side: every control step v_target = kp * (q_target - q), the heading error of a moving base wrapped to
    // (-pi, pi], then setTargetVelocity (incl. its scaleToLimits) and the velocity controller.
    bool positionController = false;
    std::vector<float> targetPosition;
    float positionGain = 0.0f;

    unsigned numBaseDofs() const { return isStatic ? 0 : 3; }
    std::vector<int> allDofs() const;
    std::vector<float> getDOFPositions(const std::vector<int>& idx) const;
    std::vector<float> getDOFVelocities(const std::vector<int>& idx) const;
    void getDofLimits(int dof, DofInfo* info) const;
    void setDOFPositions(const std::vector<float>& values, const std::vector<int>& idx);
    void setDOFVelocities(const std::vector<float>& values, const std::vector<int>& idx);
    void setPose(float x, float y, float theta);
    // Box2DObject::getBallApproximation (:1754-1763): links in std::map (name) order
    void ballApproximation(std::vector<float>& out) const;
    void updateBodyVelocities(const std::vector<float>& allDofVel);
    float getInertia() const;
    // controller
    void setTargetVelocity(const std::vector<float>& v);
    bool controllerCompute(const std::vector<float>& velocities, float dt, std::vector<float>& out) const;
    float kinematicChainInertia(const OJoint* root) const;
    void control(float dt);
    void commandEfforts(const std::vector<float>& efforts);
};

struct OContact {
    float point[3];
    float normal[3];
    int bodyA, bodyB;        // b2o body indices (fixtureA / fixtureB bodies)
    int fixtureA, fixtureB;  // proxy ids
};

struct OWorld : ContactListener {
    EnvDesc env;
    std::unique_ptr<World> b2;
    Body* ground = nullptr;
    float scale = 1.0f;
    float timeStep = 0.01f;
    int velocitySteps = 15;
    int positionSteps = 15;
    std::vector<std::unique_ptr<OObject>> owned;
    std::vector<std::unique_ptr<OLink>> ownedLinks;
    std::vector<std::unique_ptr<OJoint>> ownedJoints;
    std::map<std::string, OObject*> objects;  // name order
    std::map<std::string, OObject*> robots;   // name order
    std::vector<OObject*> creationOrder;      // robots then objects, file order
    bool recordContacts = false;
    std::vector<OContact> recorded;
    std::vector<std::string> warnings;

    explicit OWorld(const EnvDesc& e);
    float invScale() const { return 1.0f / scale; }
    float gravity() const { return 9.81f * scale; }
    OObject* find(const std::string& name);
    void stepPhysics(int steps, bool executeControllers = true, bool allowSleeping = true);
    void stepPhysicsRecord(std::vector<OContact>& contacts, int steps, bool executeControllers, bool allowSleeping);
    void setToRest();
    // checkCollision(std::vector<Contact>&) — all touching contacts after UpdateContacts()
    void updateContactCache(std::vector<OContact>& out);
    void BeginContact(Contact* c) override;
    // getObjects(aabb, ...) (:3207-3254, Box2DAABBQuery :2979-2998): hits[i] = 1 for creationOrder[i]
    void getObjectsInAABB(const float aabb[4], bool excludeRobots, std::vector<int>& hits);
    // isPhysicallyFeasible (:3370-3400); convex clipping in place of boost::geometry::intersection
    bool isPhysicallyFeasible();
    OContact makeContact(Contact* c) const;
    // full DOF state in creation order (robots, then objects): dynamic objects contribute
    // [x, y, theta, joints...], static objects [joints...]
    int stateDim() const;
    void getState(std::vector<float>& q, std::vector<float>& qd) const;
    void setState(const std::vector<float>& q, const std::vector<float>& qd);

private:
    OObject* createObject(const ObjectDesc& d);
};

}  // namespace b2o
