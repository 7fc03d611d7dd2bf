// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// sim_env wrapper semantics restated; every routine cites the reference lines it follows.
#include "b2o_simenv.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace b2o {

static inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }

// boost::geometry::area on a CCW float polygon: the surveyor strategy accumulates in double
// (select_most_precise<float,double>), Box2DWorld.cpp:101.
static double polygonAreaDouble(const std::vector<float>& poly) {
    size_t n = poly.size() / 2;
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) {
        size_t j = (i + 1) % n;
        double x1 = poly[2 * i], y1 = poly[2 * i + 1];
        double x2 = poly[2 * j], y2 = poly[2 * j + 1];
        sum += (x1 + x2) * (y1 - y2);
    }
    double a = sum / 2.0;
    return a < 0 ? -a : a;  // bg::correct fixes orientation, area is positive
}

// ================================================================================================ OLink
void OLink::getPose(float pose[3]) const {  // Box2DWorld.cpp:215-224
    Vec2 pos = body->GetWorldPoint(localOriginOffset);
    float inv = world->invScale();
    pose[0] = inv * pos.x;
    pose[1] = inv * pos.y;
    pose[2] = body->GetAngle();
}

void OLink::ballApproximation(std::vector<float>& out) const {  // :733-747 with getTransform() :193-213
    // getTransform(): matrix(0,0) = cos(a), (0,1) = -sin(a), (1,0) = -(0,1), (1,1) = (0,0), translation = invScale *
    // GetWorldPoint(_local_origin_offset).  cos / sin of the body angle are the engine's own (Rot), as everywhere else.
    Vec2 pos = body->GetWorldPoint(localOriginOffset);
    float inv = world->invScale();
    Rot r(body->GetAngle());
    const float m00 = r.c, m01 = -r.s, m10 = -m01, m11 = m00;
    const float tx = inv * pos.x, ty = inv * pos.y;
    for (const std::vector<float>& b : balls) {
        // Eigen: Affine3f * Vector3f = linear * v + translation, coefficient sums left to right
        out.push_back(((m00 * b[0] + m01 * b[1]) + 0.0f * b[2]) + tx);
        out.push_back(((m10 * b[0] + m11 * b[1]) + 0.0f * b[2]) + ty);
        out.push_back(((0.0f * b[0] + 0.0f * b[1]) + 1.0f * b[2]) + 0.0f);
        out.push_back(b[3]);
    }
}

void OLink::getVelocityVector(float v[3]) const {  // :233-241
    float inv = world->invScale();
    v[0] = inv * body->linearVelocity.x;
    v[1] = inv * body->linearVelocity.y;
    v[2] = body->angularVelocity;
}

void OLink::setPose(const float pose[3], bool updateChildren, bool jointOverride) {  // :405-438
    std::vector<float> childValues;
    if (updateChildren) {
        for (OJoint* cj : childJoints) childValues.push_back(jointOverride ? 0.0f : cj->getPosition());
    }
    float s = world->scale;
    Vec2 baseTranslation(s * pose[0], s * pose[1]);
    float theta = pose[2];
    Rot rot(theta);
    Vec2 offset = Mul(rot, localOriginOffset);
    body->SetTransform(baseTranslation - offset, theta);
    if (updateChildren) {
        size_t id = 0;
        for (OJoint* cj : childJoints) {
            cj->resetPosition(childValues.at(id), jointOverride);
            ++id;
        }
    }
}

float OLink::getInertia() const {  // :549-555
    float inv = world->invScale();
    return inv * inv * body->GetInertia();
}

float OLink::getCOMInertia() const {  // :557-566
    float inv = world->invScale();
    float inertia = body->GetInertia();
    float comDist = body->GetLocalCenter().Length();
    float comInertia = inertia - getMass() * comDist * comDist;
    return inv * inv * comInertia;
}

void OLink::getCenterOfMass(float com[2]) const {  // :595-603
    float inv = world->invScale();
    Vec2 c = body->GetWorldCenter();
    com[0] = inv * c.x;
    com[1] = inv * c.y;
}

void OLink::setOriginToCenterOfMass() {  // :653-671
    if (body->type == kStaticBody) {
        // BoundingBox::center() of the missing sim_env package: assumed (min+max)/2
        float cx = 0.5f * (localAabbMin[0] + localAabbMax[0]);
        float cy = 0.5f * (localAabbMin[1] + localAabbMax[1]);
        localOriginOffset.x = world->scale * cx;
        localOriginOffset.y = world->scale * cy;
    } else {
        localOriginOffset = body->GetLocalCenter();
    }
}

void OLink::updateBodyVelocities(const std::vector<float>& qd, std::vector<std::pair<Vec2, unsigned>>& parentJoints,
                                 unsigned indexOffset) {  // :759-794
    float s = world->scale;
    Vec2 lin(s * qd[0], s * qd[1]);
    float ang = 0.0f;
    Vec2 com = body->GetWorldCenter();
    for (size_t i = 0; i < parentJoints.size(); ++i) {
        Vec2 axis = parentJoints[i].first;
        unsigned idx = parentJoints[i].second;
        Vec2 r = com - axis;
        lin.x -= qd[idx] * r.y;
        lin.y += qd[idx] * r.x;
        ang += qd[idx];
    }
    body->SetLinearVelocity(lin);
    body->SetAngularVelocity(ang);
    for (OJoint* cj : childJoints) {
        parentJoints.emplace_back(cj->getGlobalAxisPosition(), cj->dofIndex + indexOffset);
        cj->linkB->updateBodyVelocities(qd, parentJoints, indexOffset);
        parentJoints.pop_back();
    }
}

void OLink::setMass(float mass) {  // :478-489
    MassData data;
    body->GetMassData(&data);
    float massRatio = mass / data.mass;
    data.mass = mass;
    data.I = data.I * massRatio;
    body->SetMassData(&data);
    updateFrictionJoint();
}

void OLink::updateFrictionJoint() {  // :523-531 (note the operand order: mu * mass * gravity, the ctor has mu * gravity * mass)
    float gravity = 9.81f * world->scale;
    frictionJoint->maxForce = groundFriction * getMass() * gravity;
    frictionJoint->maxTorque = groundFriction * groundTorqueIntegral * world->scale * world->scale;
}

void OLink::setContactFriction(float mu) {  // :538-547; b2Fixture::SetFriction: live contacts keep their mixed value
    contactFriction = mu;
    for (Fixture* f : body->fixtures) f->friction = mu;
}

void OLink::setEnabled(bool e) {  // :802-822
    if (e == enabled) return;
    enabled = e;
    for (int i = (int)body->fixtures.size() - 1; i >= 0; --i) {
        Fixture* f = body->fixtures[i];
        f->filter.categoryBits = e ? 1 : 0;
        // b2Fixture::SetFilterData -> Refilter: flag contacts, touch proxies
        for (Contact* c : body->contacts) {
            if (c->fixtureA == f || c->fixtureB == f) c->filterFlag = true;
        }
        body->world->moveBuffer.push_back(f->proxyId);
    }
}

// =============================================================================================== OJoint
void OJoint::resetPosition(float value, bool childJointOverride) {  // :992-1028
    float clamped = clampf(value, positionLimits[0], positionLimits[1]);
    if (std::abs(clamped - value) > 0.0001f) world->warnings.push_back("joint position out of limits: " + name);
    if (prismatic()) return;  // :1021-1026: "TODO" -- the child link is left where it is
    float newAngle = joint->bodyA->GetAngle() + clamped + joint->referenceAngle;
    Vec2 axisWorld = joint->GetAnchorA();
    float pose[3];
    pose[0] = world->invScale() * axisWorld.x;
    pose[1] = world->invScale() * axisWorld.y;
    pose[2] = newAngle;
    linkB->setPose(pose, true, childJointOverride);
}

void OJoint::setControlTorque(float value) {  // :1198-1222  (the "motor hack")
    if (prismatic()) throw std::logic_error("setControlTorque for prismatic joints is not implemented yet");  // :1204-1208
    float s = world->scale;
    joint->EnableMotor(true);
    joint->SetMaxMotorTorque(s * s * std::abs(value));
    float sgn = value > 0.0f ? 1.0f : -1.0f;
    joint->SetMotorSpeed(sgn * std::numeric_limits<float>::max());
}

void OJoint::getAxisPosition(float axis[2]) const {  // :1249-1257
    float pose[3];
    linkB->getPose(pose);
    axis[0] = pose[0];
    axis[1] = pose[1];
}

// ============================================================================================== OObject
std::vector<int> OObject::allDofs() const {
    std::vector<int> v(numDofs);
    for (unsigned i = 0; i < numDofs; ++i) v[i] = (int)i;
    return v;
}

void OObject::ballApproximation(std::vector<float>& out) const {  // :1754-1763
    for (const auto& kv : links) kv.second->ballApproximation(out);
}

std::vector<float> OObject::getDOFPositions(const std::vector<int>& idxIn) const {  // :1507-1528
    const std::vector<int>& idx = idxIn.empty() ? activeDofs : idxIn;
    std::vector<float> out(idx.size());
    for (size_t i = 0; i < idx.size(); ++i) {
        int dof = idx[i];
        if (dof < 3 && !isStatic) {
            float pose[3];
            baseLink->getPose(pose);
            out[i] = pose[dof];
        } else {
            int ji = isStatic ? dof : dof - 3;
            out[i] = sortedJoints[ji]->getPosition();
        }
    }
    return out;
}

std::vector<float> OObject::getDOFVelocities(const std::vector<int>& idxIn) const {  // :1628-1649
    const std::vector<int>& idx = idxIn.empty() ? activeDofs : idxIn;
    std::vector<float> out(idx.size());
    for (size_t i = 0; i < idx.size(); ++i) {
        int dof = idx[i];
        if (!isStatic && dof < 3) {
            float v[3];
            baseLink->getVelocityVector(v);
            out[i] = v[dof];
        } else {
            int ji = isStatic ? dof : dof - 3;
            out[i] = sortedJoints[ji]->getVelocity();
        }
    }
    return out;
}

void OObject::getDofLimits(int dof, DofInfo* info) const {  // :1557-1583, 1224-1230
    if ((unsigned)dof < numBaseDofs()) {
        *info = baseDof[dof];
    } else {
        const OJoint* j = sortedJoints[dof - numBaseDofs()];
        for (int k = 0; k < 2; ++k) {
            info->positionLimits[k] = j->positionLimits[k];
            info->velocityLimits[k] = j->velocityLimits[k];
            info->accelerationLimits[k] = j->accelerationLimits[k];
        }
    }
}

void OObject::setDOFPositions(const std::vector<float>& values, const std::vector<int>& idxIn) {  // :1585-1626
    const std::vector<int> idx = idxIn.empty() ? activeDofs : idxIn;
    if (idx.size() != values.size()) throw std::runtime_error("setDOFPositions: size mismatch");
    std::vector<float> velocities = getDOFVelocities(allDofs());
    unsigned dofOffset = 0;
    float pose[3];
    baseLink->getPose(pose);
    for (unsigned i = 0; i < std::min(numBaseDofs(), (unsigned)idx.size()); ++i) {
        unsigned d = (unsigned)idx[i];
        if (d < numBaseDofs()) {
            ++dofOffset;
            pose[d] = values[i];
        } else {
            break;
        }
    }
    if (dofOffset > 0) baseLink->setPose(pose);
    for (size_t i = dofOffset; i < idx.size(); ++i) {
        int dof = idx[i];
        int ji = isStatic ? dof : dof - 3;
        sortedJoints[ji]->resetPosition(values[i], false);
    }
    updateBodyVelocities(velocities);
}

void OObject::setDOFVelocities(const std::vector<float>& values, const std::vector<int>& idxIn) {  // :1691-1715
    const std::vector<int> idx = idxIn.empty() ? activeDofs : idxIn;
    if (idx.size() != values.size())
        throw std::runtime_error("[sim_env::Box2DObject::setDOFVelocities] Could not set DOF velocities."
                                 " Number of indices to set and number of provided values do not match.");
    std::vector<float> all = getDOFVelocities(allDofs());
    for (size_t i = 0; i < idx.size(); ++i) {
        DofInfo info;
        getDofLimits(idx[i], &info);
        all[idx[i]] = clampf(values[i], info.velocityLimits[0], info.velocityLimits[1]);
    }
    updateBodyVelocities(all);
}

void OObject::setPose(float x, float y, float theta) {  // :1940-1947 (and setTransform :1419-1433)
    std::vector<float> vel = getDOFVelocities(allDofs());
    float pose[3] = {x, y, theta};
    baseLink->setPose(pose, true);
    updateBodyVelocities(vel);
}

void OObject::updateBodyVelocities(const std::vector<float>& allDofVel) {  // :2018-2041
    std::vector<float> velocities = allDofVel;
    unsigned indexOffset = 0;
    if (isStatic) {
        velocities.assign(allDofVel.size() + 3, 0.0f);
        for (size_t i = 3; i < velocities.size(); ++i) velocities[i] = allDofVel[i - 3];
        indexOffset = 3;
    }
    std::vector<std::pair<Vec2, unsigned>> parentJoints;
    float pose[3];
    baseLink->getPose(pose);
    float s = world->scale;
    parentJoints.emplace_back(Vec2(s * pose[0], s * pose[1]), 2u);
    baseLink->updateBodyVelocities(velocities, parentJoints, indexOffset);
}

float OObject::getInertia() const {  // :1909-1921
    float com[2];
    baseLink->getCenterOfMass(com);
    float inertia = 0.0f;
    for (auto& kv : links) {
        OLink* link = kv.second;
        float lc[2];
        link->getCenterOfMass(lc);
        float dx = lc[0] - com[0], dy = lc[1] - com[1];
        float distance = sqrtf(dx * dx + dy * dy);
        inertia += link->getCOMInertia() + distance * distance * link->getMass();
    }
    return inertia;
}

// ---- Box2DRobotVelocityController ------------------------------------------------------------------
void OObject::setTargetVelocity(const std::vector<float>& v) {  // Box2DController.cpp:42-63
    if (v.size() != activeDofs.size())
        throw std::runtime_error("[sim_env::Box2DRobotVelocityController::setTargetVelocity]"
                                 "Provided target velocity has invalid dimension.");
    targetVelocity = v;
    // utils::eigen::scaleToLimits lives in the missing sim_env package; assumed: one uniform scale
    // factor that brings every component inside its velocity limits.
    float factor = 1.0f;
    for (size_t i = 0; i < v.size(); ++i) {
        DofInfo info;
        getDofLimits(activeDofs[i], &info);
        if (v[i] > info.velocityLimits[1] && v[i] > 0.0f) factor = std::min(factor, info.velocityLimits[1] / v[i]);
        if (v[i] < info.velocityLimits[0] && v[i] < 0.0f) factor = std::min(factor, info.velocityLimits[0] / v[i]);
    }
    if (factor < 1.0f)
        for (float& x : targetVelocity) x = factor * x;
    targetAvailable = true;
}

float OObject::kinematicChainInertia(const OJoint* root) const {  // Box2DController.cpp:191-215
    float axis[2];
    root->getAxisPosition(axis);
    float inertia = 0.0f;
    // the reference's queue never receives the child joints: only the root's child link counts
    OLink* child = root->linkB;
    float com[2];
    child->getCenterOfMass(com);
    float dx = com[0] - axis[0], dy = com[1] - axis[1];
    float distance = sqrtf(dx * dx + dy * dy);
    inertia += child->getInertia() + child->getMass() * distance * distance;
    return inertia;
}

bool OObject::controllerCompute(const std::vector<float>& velocities, float dt, std::vector<float>& out) const {
    // Box2DController.cpp:65-117
    if (!targetAvailable) return false;
    out.resize(activeDofs.size());
    for (size_t i = 0; i < activeDofs.size(); ++i) {
        float massInertia = 0.0f;
        int dof = activeDofs[i];
        if ((unsigned)dof < numBaseDofs()) {
            if (dof == 0 || dof == 1) massInertia = mass;
            else massInertia = getInertia();
        } else {
            massInertia = kinematicChainInertia(sortedJoints[dof - numBaseDofs()]);
        }
        DofInfo info;
        getDofLimits(dof, &info);
        float accel = (targetVelocity[i] - velocities[i]) / dt;
        accel = std::min(info.accelerationLimits[1], std::max(accel, info.accelerationLimits[0]));
        out[i] = massInertia * accel;
    }
    return true;
}

void OObject::control(float dt) {  // Box2DWorld.cpp:2284-2310
    if (!isRobot) return;
    if (effortController) {  // callback success with these efforts
        commandEfforts(controlEfforts);
        return;
    }
    if (positionController) {
        std::vector<float> q = getDOFPositions({});
        std::vector<float> v(q.size());
        for (size_t i = 0; i < q.size(); ++i) {
            float err = targetPosition[i] - q[i];
            if (!isStatic && activeDofs[i] == 2) err = err - 6.2831855f * rintf(err * 0.15915494f);
            v[i] = positionGain * err;
        }
        setTargetVelocity(v);
    }
    std::vector<float> velocities = getDOFVelocities({});
    std::vector<float> efforts;
    if (controllerCompute(velocities, dt, efforts)) commandEfforts(efforts);
}

void OObject::commandEfforts(const std::vector<float>& target) {  // :2319-2379
    float s = world->scale;
    for (size_t i = 0; i < activeDofs.size(); ++i) {
        int dof = activeDofs[i];
        if ((unsigned)dof < numBaseDofs()) {
            Body* body = baseLink->body;
            switch (dof) {
            case 0: body->ApplyForceToCenter(Vec2(s * target[i], 0.0f), true); break;
            case 1: body->ApplyForceToCenter(Vec2(0.0f, s * target[i]), true); break;
            case 2: body->ApplyTorque(s * s * target[i], true); break;
            }
        } else {
            sortedJoints[dof - numBaseDofs()]->setControlTorque(target[i]);
        }
    }
}

// =============================================================================================== OWorld
OObject* OWorld::createObject(const ObjectDesc& d) {  // Box2DObject ctor :1292-1363, Box2DRobot ctor :2046-2067
    owned.emplace_back(new OObject());
    OObject* obj = owned.back().get();
    obj->name = d.name;
    obj->isStatic = d.is_static;
    obj->isRobot = d.is_robot;
    obj->world = this;
    obj->mass = 0.0f;
    for (const LinkDesc& ld : d.links) {  // Box2DLink ctor :44-148
        ownedLinks.emplace_back(new OLink());
        OLink* link = ownedLinks.back().get();
        link->name = ld.name;
        link->world = this;
        link->object = obj;
        link->localOriginOffset = Vec2(0.0f, 0.0f);
        link->balls = ld.balls;  // :54-57
        link->body = b2->CreateBody(d.is_static ? kStaticBody : kDynamicBody, Vec2(0.0f, 0.0f), 0.0f);
        link->body->userData = link;
        float area = 0.0f;
        std::vector<PolygonShape> shapes;
        link->localAabbMin[0] = link->localAabbMin[1] = std::numeric_limits<float>::max();
        link->localAabbMax[0] = link->localAabbMax[1] = std::numeric_limits<float>::lowest();
        for (const std::vector<float>& poly : ld.polygons) {
            std::vector<Vec2> pts;
            for (size_t i = 0; i < poly.size() / 2; ++i) {
                link->localAabbMin[0] = std::min(poly[2 * i], link->localAabbMin[0]);
                link->localAabbMin[1] = std::min(poly[2 * i + 1], link->localAabbMin[1]);
                link->localAabbMax[0] = std::max(poly[2 * i], link->localAabbMax[0]);
                link->localAabbMax[1] = std::max(poly[2 * i + 1], link->localAabbMax[1]);
                pts.emplace_back(scale * poly[2 * i], scale * poly[2 * i + 1]);
            }
            area += scale * scale * polygonAreaDouble(poly);  // float*float*double, accumulated to float
            PolygonShape shape;
            shape.Set(pts.data(), (int)pts.size());
            shapes.push_back(shape);
        }
        for (const PolygonShape& shape : shapes)
            link->body->CreateFixture(shape, ld.mass / area, ld.contact_friction, ld.restitution, false);
        link->contactFriction = ld.contact_friction;
        link->groundTorqueIntegral = ld.ground_torque_integral;
        link->groundFriction = ld.ground_friction;
        float g = gravity();
        float maxForce = link->groundFriction * g * link->body->GetMass();
        float maxTorque = link->groundFriction * link->groundTorqueIntegral * scale * scale;
        link->frictionJoint =
            b2->CreateFrictionJoint(ground, link->body, Vec2(0.0f, 0.0f), link->body->GetLocalCenter(), maxForce, maxTorque);
        obj->links[link->name] = link;
        obj->linksCreation.push_back(link);
        obj->mass += link->getMass();
    }
    obj->baseLink = obj->links.at(d.base_link);
    {   // :1304-1312: merge the link boxes (BoundingBox::merge of the absent sim_env: assumed component-wise min / max from an
        // empty box), then shift by the base link's local centre of mass as it is before the origin moves there
        obj->localAabb[0] = obj->localAabb[1] = std::numeric_limits<float>::max();
        obj->localAabb[2] = obj->localAabb[3] = std::numeric_limits<float>::lowest();
        for (OLink* l : obj->linksCreation) {
            obj->localAabb[0] = std::min(obj->localAabb[0], l->localAabbMin[0]);
            obj->localAabb[1] = std::min(obj->localAabb[1], l->localAabbMin[1]);
            obj->localAabb[2] = std::max(obj->localAabb[2], l->localAabbMax[0]);
            obj->localAabb[3] = std::max(obj->localAabb[3], l->localAabbMax[1]);
        }
        Vec2 b2center = obj->baseLink->body->GetLocalCenter() - obj->baseLink->localOriginOffset;  // getLocalCenterOfMass :577-586
        float oldCenter[2] = {b2center.x * invScale(), b2center.y * invScale()};
        obj->localAabb[0] -= oldCenter[0]; obj->localAabb[1] -= oldCenter[1];
        obj->localAabb[2] -= oldCenter[0]; obj->localAabb[3] -= oldCenter[1];
    }
    obj->baseLink->setOriginToCenterOfMass();
    unsigned baseDofs = obj->isStatic ? 0 : 3;
    for (const JointDesc& jd : d.joints) {  // Box2DJoint ctor :872-948
        // unknown types fall back to revolute with an error message (:893-898)
        const bool prismatic = jd.joint_type == "prismatic";
        if (!prismatic && jd.joint_type != "revolute") warnings.push_back("Unknown joint type " + jd.joint_type + ". Using revolute instead.");
        ownedJoints.emplace_back(new OJoint());
        OJoint* j = ownedJoints.back().get();
        j->name = jd.name;
        j->world = this;
        j->object = obj;
        j->linkA = obj->links.at(jd.link_a);
        j->linkB = obj->links.at(jd.link_b);
        Vec2 axis(scale * jd.axis[0], scale * jd.axis[1]);
        float norm = sqrtf(jd.position_limits[0] * jd.position_limits[0] + jd.position_limits[1] * jd.position_limits[1]);
        bool enableLimit = norm > 0.0;
        if (prismatic) {
            // :920-940 -- localAnchorA is NOT scaled there, the translation limits are; the axis direction is the unit vector
            // at axis_orientation (libm cos / sin of a description constant); maxMotorForce = min |acceleration limits|
            Vec2 anchorA(jd.axis[0], jd.axis[1]);
            Vec2 direction(std::cos(jd.axis_orientation), std::sin(jd.axis_orientation));
            float maxMotorForce = std::min(std::abs(jd.acceleration_limits[0]), std::abs(jd.acceleration_limits[1]));
            j->joint = b2->CreatePrismaticJoint(j->linkA->body, j->linkB->body, anchorA, Vec2(0.0f, 0.0f), direction, jd.axis_orientation,
                                                scale * jd.position_limits[0], scale * jd.position_limits[1], enableLimit, maxMotorForce,
                                                jd.actuated);
        } else {
            j->joint = b2->CreateRevoluteJoint(j->linkA->body, j->linkB->body, axis, Vec2(0.0f, 0.0f), jd.axis_orientation,
                                               jd.position_limits[0], jd.position_limits[1], enableLimit, jd.actuated);
        }
        for (int k = 0; k < 2; ++k) {
            j->positionLimits[k] = jd.position_limits[k];
            j->velocityLimits[k] = jd.velocity_limits[k];
            j->accelerationLimits[k] = jd.acceleration_limits[k];
        }
        j->linkA->childJoints.push_back(j);
        j->linkB->parentJoint = j;
        obj->sortedJoints.push_back(j);
        j->jointIndex = (unsigned)obj->sortedJoints.size() - 1;
        j->dofIndex = j->jointIndex + baseDofs;
    }
    obj->numDofs = baseDofs + (unsigned)obj->sortedJoints.size();
    obj->activeDofs = obj->allDofs();
    float zero[3] = {0.0f, 0.0f, 0.0f};
    obj->baseLink->setPose(zero, true, true);
    const float lo = std::numeric_limits<float>::lowest(), hi = std::numeric_limits<float>::max();
    for (int k = 0; k < 3; ++k) {
        obj->baseDof[k].positionLimits[0] = lo; obj->baseDof[k].positionLimits[1] = hi;
        obj->baseDof[k].velocityLimits[0] = lo; obj->baseDof[k].velocityLimits[1] = hi;
        obj->baseDof[k].accelerationLimits[0] = lo; obj->baseDof[k].accelerationLimits[1] = hi;
    }
    obj->baseDof[2].positionLimits[0] = (float)-M_PI;
    obj->baseDof[2].positionLimits[1] = (float)M_PI;
    if (d.is_robot) {
        if (d.use_center_of_mass) obj->baseLink->setOriginToCenterOfMass();
        if (!obj->isStatic) {
            for (int k = 0; k < 2; ++k) {
                obj->baseDof[0].velocityLimits[k] = d.translational_velocity_limits[k];
                obj->baseDof[0].accelerationLimits[k] = d.translational_acceleration_limits[k];
                obj->baseDof[1].velocityLimits[k] = d.translational_velocity_limits[k];
                obj->baseDof[1].accelerationLimits[k] = d.translational_acceleration_limits[k];
                obj->baseDof[2].velocityLimits[k] = d.rotational_velocity_limits[k];
                obj->baseDof[2].accelerationLimits[k] = d.rotational_acceleration_limits[k];
            }
        }
    }
    creationOrder.push_back(obj);
    return obj;
}

OWorld::OWorld(const EnvDesc& e) : env(e) {  // createWorld :3664-3721, createGround :3644-3662
    b2.reset(new World(Vec2(0.0f, 0.0f)));
    scale = e.scale;
    float wb[4] = {e.world_bounds[0], e.world_bounds[1], e.world_bounds[2], e.world_bounds[3]};
    float n = sqrtf(wb[0] * wb[0] + wb[1] * wb[1] + wb[2] * wb[2] + wb[3] * wb[3]);
    if (n == 0.0) {
        wb[0] = scale * -100.0f; wb[1] = scale * -100.0f; wb[2] = scale * 100.0f; wb[3] = scale * 100.0f;
    }
    ground = b2->CreateBody(kStaticBody, Vec2(0.5f * scale * (wb[0] + wb[2]), 0.5f * scale * (wb[1] + wb[3])), 0.0f);
    PolygonShape groundShape;
    groundShape.SetAsBox(scale * (wb[2] - wb[0]), scale * (wb[3] - wb[1]));
    ground->CreateFixture(groundShape, 0.0f, 0.2f, 0.0f, true);

    for (const ObjectDesc& d : e.robots) {
        ObjectDesc dd = d;
        dd.is_robot = true;
        OObject* o = createObject(dd);
        if (robots.count(o->name) || objects.count(o->name)) throw std::runtime_error("duplicate name " + o->name);
        robots[o->name] = o;
    }
    for (const ObjectDesc& d : e.objects) {
        ObjectDesc dd = d;
        dd.is_robot = false;
        OObject* o = createObject(dd);
        if (robots.count(o->name) || objects.count(o->name)) throw std::runtime_error("duplicate name " + o->name);
        objects[o->name] = o;
    }
    for (auto& kv : e.states) {  // :3684-3717 (std::map: name order)
        OObject* o = find(kv.first);
        if (!o) continue;
        const StateDesc& sd = kv.second;
        if (!o->isStatic) {
            o->setDOFPositions(sd.configuration, {});
            o->setDOFVelocities(sd.velocity, {});
        } else {
            o->setPose(sd.configuration.at(0), sd.configuration.at(1), sd.configuration.at(2));
            std::vector<float> rc, rv;
            for (size_t i = 3; i < sd.configuration.size(); ++i) {
                rc.push_back(sd.configuration[i]);
                rv.push_back(sd.velocity.at(i));
            }
            o->setDOFPositions(rc, {});
            o->setDOFVelocities(rv, {});
        }
    }
    b2->listener = this;
}

OObject* OWorld::find(const std::string& name) {
    auto it = objects.find(name);
    if (it != objects.end()) return it->second;
    it = robots.find(name);
    if (it != robots.end()) return it->second;
    return nullptr;
}

void OWorld::stepPhysics(int steps, bool executeControllers, bool allowSleeping) {  // :3266-3283
    for (int i = 0; i < steps; ++i) {
        if (executeControllers) {
            for (auto& kv : robots) kv.second->control(timeStep);
        }
        b2->SetAllowSleeping(allowSleeping);
        b2->Step(timeStep, velocitySteps, positionSteps);
        b2->ClearForces();
        b2->SetAllowSleeping(true);
    }
}

void OWorld::setToRest() {  // :3550-3558, :1717-1722
    for (auto& kv : objects) kv.second->setDOFVelocities(std::vector<float>(kv.second->numDofs, 0.0f), kv.second->allDofs());
    for (auto& kv : robots) kv.second->setDOFVelocities(std::vector<float>(kv.second->numDofs, 0.0f), kv.second->allDofs());
}

OContact OWorld::makeContact(Contact* c) const {  // :2797-2807, :2853-2863
    WorldManifold wm;
    wm.points[0] = Vec2(0.0f, 0.0f);
    c->GetWorldManifold(&wm);
    OContact oc;
    float inv = invScale();
    oc.point[0] = inv * wm.points[0].x;
    oc.point[1] = inv * wm.points[0].y;
    oc.point[2] = 0.0f;
    oc.normal[0] = inv * wm.normal.x;
    oc.normal[1] = inv * wm.normal.y;
    oc.normal[2] = 0.0f;
    oc.bodyA = c->fixtureA->body->index;
    oc.bodyB = c->fixtureB->body->index;
    oc.fixtureA = c->fixtureA->proxyId;
    oc.fixtureB = c->fixtureB->proxyId;
    return oc;
}

void OWorld::BeginContact(Contact* c) {  // :2792-2818
    if (recordContacts) recorded.push_back(makeContact(c));
}

void OWorld::updateContactCache(std::vector<OContact>& out) {  // :2880-2911
    b2->UpdateContacts();
    for (int i = (int)b2->contacts.size() - 1; i >= 0; --i) {
        Contact* c = b2->contacts[i];
        if (c->IsTouching() && c->IsEnabled()) out.push_back(makeContact(c));
    }
}

void OWorld::stepPhysicsRecord(std::vector<OContact>& contacts, int steps, bool executeControllers, bool allowSleeping) {
    // :3290-3310 with startRecordings :2743-2768
    recorded.clear();
    recordContacts = false;
    std::vector<float> q, qd;
    getState(q, qd);
    setToRest();
    stepPhysics(1, false, false);
    recordContacts = true;
    for (int i = (int)b2->contacts.size() - 1; i >= 0; --i) {
        Contact* c = b2->contacts[i];
        if (c->IsTouching() && c->IsEnabled()) BeginContact(c);
    }
    setState(q, qd);
    stepPhysics(steps, executeControllers, allowSleeping);
    recordContacts = false;
    contacts = recorded;
}

void OWorld::getObjectsInAABB(const float aabb[4], bool excludeRobots, std::vector<int>& hits) {
    hits.assign(creationOrder.size(), 0);
    AABB query;
    query.lowerBound = Vec2(scale * aabb[0], scale * aabb[1]);  // :3210-3214
    query.upperBound = Vec2(scale * aabb[2], scale * aabb[3]);
    Vec2 extents = 0.5f * (query.upperBound - query.lowerBound);  // b2AABB::GetExtents
    Vec2 center = 0.5f * (query.lowerBound + query.upperBound);   // b2AABB::GetCenter
    PolygonShape box;
    box.SetAsBox(extents.x, extents.y);
    Transform xfQ;
    xfQ.p = center;
    xfQ.q.s = 0.0f;  // b2Transform::Set(center, 0)
    xfQ.q.c = 1.0f;
    // b2World::QueryAABB: every proxy whose fat AABB overlaps the box is reported
    for (Body* body : b2->bodies) {
        if (body == ground) continue;  // b_collision_ignore (:2988)
        OLink* link = (OLink*)body->userData;
        for (Fixture* f : body->fixtures) {
            if (!TestOverlap(query, f->fatAABB)) continue;
            if (!TestOverlapShapes(&box, &f->shape, xfQ, body->xf)) continue;
            for (size_t i = 0; i < creationOrder.size(); ++i)
                if (creationOrder[i] == link->object && !(excludeRobots && link->object->isRobot)) hits[i] = 1;
        }
    }
}

namespace {
// Sutherland-Hodgman: area of the intersection of two convex counter-clockwise polygons
float convexIntersectionArea(const Vec2* a, int na, const Vec2* b, int nb) {
    Vec2 bufA[18], bufB[18];
    Vec2* in = bufA;
    Vec2* out = bufB;
    int n = na;
    for (int i = 0; i < na; ++i) in[i] = a[i];
    for (int e = 0; e < nb && n > 0; ++e) {
        const Vec2 c1 = b[e], c2 = b[e + 1 < nb ? e + 1 : 0];
        const Vec2 edge = c2 - c1;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const Vec2 p = in[i], q = in[i + 1 < n ? i + 1 : 0];
            const float dp = Cross(edge, p - c1), dq = Cross(edge, q - c1);
            if (dq >= 0.0f) {
                if (dp < 0.0f) { float t = dp / (dp - dq); out[m++] = p + t * (q - p); }
                out[m++] = q;
            } else if (dp >= 0.0f) {
                float t = dp / (dp - dq);
                out[m++] = p + t * (q - p);
            }
        }
        Vec2* tmp = in; in = out; out = tmp;
        n = m;
    }
    if (n < 3) return 0.0f;
    float twice = 0.0f;
    for (int i = 1; i + 1 < n; ++i) twice += Cross(in[i] - in[0], in[i + 1] - in[0]);
    return 0.5f * twice;
}
}  // namespace

bool OWorld::isPhysicallyFeasible() {
    b2->UpdateContacts();  // checkCollision(contacts) :3373
    const float inv = invScale();
    for (int k = (int)b2->contacts.size() - 1; k >= 0; --k) {
        Contact* c = b2->contacts[k];
        if (!(c->IsTouching() && c->IsEnabled())) continue;
        Body* bA = c->fixtureA->body;
        Body* bB = c->fixtureB->body;
        for (Fixture* fa : bA->fixtures) {
            Vec2 va[kMaxPolygonVertices];
            for (int i = 0; i < fa->shape.count; ++i) va[i] = Mul(bA->xf, fa->shape.vertices[i]);
            for (Fixture* fb : bB->fixtures) {
                Vec2 vb[kMaxPolygonVertices];
                for (int i = 0; i < fb->shape.count; ++i) vb[i] = Mul(bB->xf, fb->shape.vertices[i]);
                float area = convexIntersectionArea(va, fa->shape.count, vb, fb->shape.count);
                if (inv * inv * area > 0.0001f) return false;  // MAX_ALLOWED_INTERSECTION_AREA (:22)
            }
        }
    }
    return true;
}

int OWorld::stateDim() const {
    int n = 0;
    for (OObject* o : creationOrder) n += (int)o->numDofs;
    return n;
}

void OWorld::getState(std::vector<float>& q, std::vector<float>& qd) const {  // getWorldState :3462-3474
    q.clear();
    qd.clear();
    for (OObject* o : creationOrder) {
        std::vector<float> p = o->getDOFPositions(o->allDofs());
        std::vector<float> v = o->getDOFVelocities(o->allDofs());
        q.insert(q.end(), p.begin(), p.end());
        qd.insert(qd.end(), v.begin(), v.end());
    }
}

void OWorld::setState(const std::vector<float>& q, const std::vector<float>& qd) {  // setWorldState :3492-3513 / setState :1980-1989
    size_t off = 0;
    for (OObject* o : creationOrder) {
        std::vector<float> p(q.begin() + off, q.begin() + off + o->numDofs);
        std::vector<float> v(qd.begin() + off, qd.begin() + off + o->numDofs);
        off += o->numDofs;
        o->activeDofs = o->allDofs();
        // setTransform(pose): the pose of a dynamic object equals its first three DOFs; the yaw is taken
        // directly instead of through acos(R00) (:1424-1425).  Static objects keep their pose.
        if (!o->isStatic) o->setPose(p[0], p[1], p[2]);
        o->setDOFPositions(p, o->allDofs());
        o->setDOFVelocities(v, o->allDofs());
    }
}

}  // namespace b2o
