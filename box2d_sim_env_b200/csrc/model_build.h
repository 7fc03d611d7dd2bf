// Host model builder: EnvironmentDescription -> flattened model (setup time, CPU).
#pragma once
#include <string>
#include <vector>

#include "desc.h"
#include "model.h"

namespace b2g {

struct HostObjectInfo {
    std::string name;
    bool isRobot = false, isStatic = false;
    int nDofs = 0, dofOffset = 0, firstBody = 0, nLinks = 0, nJoints = 0;
    int ballStart = 0, nBalls = 0;  // rows of HostModel::balls
    // the object has a prismatic joint: the reference cannot place it or set its velocities (Box2DJoint::resetPosition is a
    // no-op, getGlobalAxisPosition throws inside updateBodyVelocities; Box2DWorld.cpp:1021-1026, 1274-1287, 789)
    bool hasPrismatic = false;
};

struct HostModel {
    ModelHeader h;
    std::vector<uint32_t> blob;
    std::vector<HostObjectInfo> objects;
    std::vector<std::string> bodyNames;   // "<object>/<link>"; body 0 is the ground
    std::vector<std::string> jointNames;  // revolute joints: YAML name; ground-friction joints: ""
    // introspection copies (parity probes)
    std::vector<float> bodyModel;     // nb x 8
    std::vector<int32_t> fixInts;     // nf x 3
    std::vector<float> fixFloats;     // nf x 38
    std::vector<int32_t> jointInts;   // nj x 6
    std::vector<float> jointFloats;   // nj x 15
    std::vector<float> dofLimits;     // nq x 6
    // ball approximation (Box2DLink::_balls), object by object, links in name order (Box2DObject::getBallApproximation
    // walks its std::map of links, Box2DWorld.cpp:1754-1763): rows of 8 = body, origin offset x y (scaled), centre x y z, radius
    std::vector<float> balls;
    std::vector<std::string> warnings;
    // Box2DLink's tunable dynamics parameters per body: _ground_friction, _ground_torque_integral, _contact_friction, index
    // of the link's ground-friction joint (-1: the body is not a link, i.e. the ground)
    std::vector<float> linkParams;    // nb x 4
    std::vector<float> linkAabb;      // nb x 4: Box2DLink::_local_aabb (Box2DWorld.cpp:69-90), user units
};
void objectLocalAabb(const HostModel& m, int object, float out[4]);

// Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction and the
// matching getters (Box2DWorld.cpp:471-566), numbered as in include/b2g.h (b2g_link_property)
enum LinkProperty {
    LP_MASS = 0, LP_GROUND_FRICTION, LP_GROUND_TORQUE_INTEGRAL, LP_CONTACT_FRICTION,  // settable
    LP_INERTIA, LP_COM_INERTIA, LP_LIMIT_FORCE, LP_LIMIT_TORQUE,                      // derived, read-only
    LP_COUNT
};
bool isLinkBody(const HostModel& m, int body);
// patches blob + introspection copies in place; returns true when the per-world sweep centres of `body` have to be
// re-derived (b2Body::SetMassData on a dynamic body); throws std::runtime_error for a read-only property
bool setLinkProperty(HostModel& m, int body, int prop, float value);
float getLinkProperty(const HostModel& m, int body, int prop);

// throws std::runtime_error (unsupported joint type, capacity, malformed description)
void buildModel(const EnvironmentDescription& env, int maxContacts, HostModel& out);
// recompute the per-world state layout for a different contact capacity
void layoutState(ModelHeader& h, int maxContacts);

}  // namespace b2g
