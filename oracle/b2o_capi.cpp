// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// Plain C entry points so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs can drive the oracle through ctypes.  Nothing in the product links this.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "b2o_simenv.h"

using namespace b2o;

namespace {
thread_local std::string g_err;
ObjectDesc& pick(EnvDesc* e, int isRobot, int idx) { return isRobot ? e->robots.at(idx) : e->objects.at(idx); }

void finalizeEnv(EnvDesc& e) {
    // base link = first link without a parent joint (Box2DIOUtils.h:212-257 verifyKinematicStructure)
    auto fix = [](ObjectDesc& o) {
        if (!o.base_link.empty()) return;
        std::map<std::string, int> parents;
        for (auto& l : o.links) parents[l.name] = 0;
        for (auto& j : o.joints) parents[j.link_b] += 1;
        for (auto& l : o.links) {
            if (parents[l.name] == 0) { o.base_link = l.name; break; }
        }
    };
    for (auto& o : e.robots) fix(o);
    for (auto& o : e.objects) fix(o);
}
}  // namespace

extern "C" {

const char* b2o_last_error() { return g_err.c_str(); }

void* b2o_env_new(float scale, const float* bounds4) {
    EnvDesc* e = new EnvDesc();
    e->scale = scale;
    for (int i = 0; i < 4; ++i) e->world_bounds[i] = bounds4 ? bounds4[i] : 0.0f;
    return e;
}
void b2o_env_free(void* env) { delete (EnvDesc*)env; }

int b2o_env_add_object(void* env, const char* name, int isRobot, int isStatic, const float* baseLimits8, int useCom) {
    EnvDesc* e = (EnvDesc*)env;
    ObjectDesc o;
    o.name = name;
    o.is_robot = isRobot != 0;
    o.is_static = isStatic != 0;
    o.use_center_of_mass = useCom != 0;
    if (baseLimits8) {
        for (int k = 0; k < 2; ++k) {
            o.translational_velocity_limits[k] = baseLimits8[k];
            o.translational_acceleration_limits[k] = baseLimits8[2 + k];
            o.rotational_velocity_limits[k] = baseLimits8[4 + k];
            o.rotational_acceleration_limits[k] = baseLimits8[6 + k];
        }
    }
    auto& vec = isRobot ? e->robots : e->objects;
    vec.push_back(o);
    return (int)vec.size() - 1;
}

int b2o_env_add_link(void* env, int isRobot, int obj, const char* name, float mass, float groundFriction,
                     float groundTorqueIntegral, float contactFriction, float restitution) {
    ObjectDesc& o = pick((EnvDesc*)env, isRobot, obj);
    LinkDesc l;
    l.name = name;
    l.mass = mass;
    l.ground_friction = groundFriction;
    l.ground_torque_integral = groundTorqueIntegral;
    l.contact_friction = contactFriction;
    l.restitution = restitution;
    o.links.push_back(l);
    return (int)o.links.size() - 1;
}

void b2o_env_add_polygon(void* env, int isRobot, int obj, int link, const float* xy, int nFloats) {
    ObjectDesc& o = pick((EnvDesc*)env, isRobot, obj);
    o.links.at(link).polygons.emplace_back(xy, xy + nFloats);
}

void b2o_env_add_ball(void* env, int isRobot, int obj, int link, const float* center3, float radius) {
    ObjectDesc& o = pick((EnvDesc*)env, isRobot, obj);
    o.links.at(link).balls.push_back({center3[0], center3[1], center3[2], radius});
}

// params: axis(2), position_limits(2), velocity_limits(2), acceleration_limits(2), axis_orientation
void b2o_env_add_joint(void* env, int isRobot, int obj, const char* name, const char* linkA, const char* linkB,
                       const float* params9, int actuated, const char* type) {
    ObjectDesc& o = pick((EnvDesc*)env, isRobot, obj);
    JointDesc j;
    j.name = name;
    j.link_a = linkA;
    j.link_b = linkB;
    j.axis[0] = params9[0]; j.axis[1] = params9[1];
    j.position_limits[0] = params9[2]; j.position_limits[1] = params9[3];
    j.velocity_limits[0] = params9[4]; j.velocity_limits[1] = params9[5];
    j.acceleration_limits[0] = params9[6]; j.acceleration_limits[1] = params9[7];
    j.axis_orientation = params9[8];
    j.actuated = actuated != 0;
    j.joint_type = type;
    o.joints.push_back(j);
}

void b2o_env_add_state(void* env, const char* name, const float* conf, const float* vel, int n) {
    EnvDesc* e = (EnvDesc*)env;
    StateDesc s;
    s.configuration.assign(conf, conf + n);
    s.velocity.assign(vel, vel + n);
    e->states[name] = s;
}

void* b2o_world_new(void* env) {
    try {
        EnvDesc e = *(EnvDesc*)env;
        finalizeEnv(e);
        return new OWorld(e);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return nullptr;
    }
}
void b2o_world_free(void* w) { delete (OWorld*)w; }

// info: [nBodies, nProxies, nJoints, stateDim, nObjects(incl robots), nContacts]
void b2o_world_info(void* wp, int* info6) {
    OWorld* w = (OWorld*)wp;
    info6[0] = (int)w->b2->bodies.size();
    info6[1] = (int)w->b2->proxies.size();
    info6[2] = (int)w->b2->joints.size();
    info6[3] = w->stateDim();
    info6[4] = (int)w->creationOrder.size();
    info6[5] = (int)w->b2->contacts.size();
}

void b2o_world_set_params(void* wp, float dt, int vit, int pit) {
    OWorld* w = (OWorld*)wp;
    w->timeStep = dt;
    w->velocitySteps = vit;
    w->positionSteps = pit;
}

void b2o_world_set_continuous(void* wp, int flag) { ((OWorld*)wp)->b2->continuousPhysics = flag != 0; }

int b2o_world_step(void* wp, int steps, int execCtrl, int allowSleep) {
    try {
        ((OWorld*)wp)->stepPhysics(steps, execCtrl != 0, allowSleep != 0);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
}

void b2o_world_get_state(void* wp, float* q, float* qd) {
    std::vector<float> a, b;
    ((OWorld*)wp)->getState(a, b);
    memcpy(q, a.data(), a.size() * sizeof(float));
    memcpy(qd, b.data(), b.size() * sizeof(float));
}

int b2o_world_set_state(void* wp, const float* q, const float* qd) {
    OWorld* w = (OWorld*)wp;
    try {
        int n = w->stateDim();
        w->setState(std::vector<float>(q, q + n), std::vector<float>(qd, qd + n));
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
}

int b2o_world_set_object_pose(void* wp, int objIdx, float x, float y, float theta) {
    OWorld* w = (OWorld*)wp;
    w->creationOrder.at(objIdx)->setPose(x, y, theta);
    return 0;
}

void b2o_world_get_object_pose(void* wp, int objIdx, float* pose3) {
    ((OWorld*)wp)->creationOrder.at(objIdx)->baseLink->getPose(pose3);
}

// Box2DObject::setDOFPositions / setDOFVelocities(values, indices); n == 0: all DOFs
int b2o_world_set_object_dofs(void* wp, int objIdx, int velocities, const int* idx, int n, const float* values) {
    OObject* o = ((OWorld*)wp)->creationOrder.at(objIdx);
    try {
        std::vector<int> ix = n > 0 ? std::vector<int>(idx, idx + n) : o->allDofs();
        std::vector<float> v(values, values + ix.size());
        if (velocities) o->setDOFVelocities(v, ix); else o->setDOFPositions(v, ix);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
}
void b2o_world_get_object_dofs(void* wp, int objIdx, int velocities, const int* idx, int n, float* out) {
    OObject* o = ((OWorld*)wp)->creationOrder.at(objIdx);
    std::vector<int> ix = n > 0 ? std::vector<int>(idx, idx + n) : o->allDofs();
    std::vector<float> v = velocities ? o->getDOFVelocities(ix) : o->getDOFPositions(ix);
    memcpy(out, v.data(), v.size() * sizeof(float));
}

// Box2DObject::getBallApproximation: returns the number of balls; out (may be null) receives [n][4]
int b2o_world_ball_approximation(void* wp, int objIdx, float* out) {
    std::vector<float> v;
    ((OWorld*)wp)->creationOrder.at(objIdx)->ballApproximation(v);
    if (out) memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size() / 4;
}

int b2o_world_set_control_efforts(void* wp, int objIdx, const float* e, int n) {
    OObject* o = ((OWorld*)wp)->creationOrder.at(objIdx);
    if (!o->isRobot || (size_t)n != o->activeDofs.size()) { g_err = "set_control_efforts: not a robot / wrong size"; return 1; }
    o->controlEfforts.assign(e, e + n);
    o->effortController = true;
    return 0;
}

int b2o_world_set_target_velocity(void* wp, int objIdx, const float* v, int n) {
    OWorld* w = (OWorld*)wp;
    try {
        w->creationOrder.at(objIdx)->effortController = false;
        w->creationOrder.at(objIdx)->positionController = false;
        w->creationOrder.at(objIdx)->setTargetVelocity(std::vector<float>(v, v + n));
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
}

// position controller (assumed law, b2o_simenv.h): target positions for all DOFs of robot objIdx and the gain
int b2o_world_set_target_position(void* wp, int objIdx, const float* q, int n, float kp) {
    OWorld* w = (OWorld*)wp;
    try {
        OObject* o = w->creationOrder.at(objIdx);
        if (!o->isRobot || (size_t)n != o->activeDofs.size()) throw std::runtime_error("bad position target");
        o->effortController = false;
        o->positionController = true;
        o->targetPosition.assign(q, q + n);
        o->positionGain = kp;
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
}

// per object: [isRobot, isStatic, numDofs, numLinks, numJoints, firstBodyIndex]
void b2o_world_object_info(void* wp, int objIdx, int* info6, char* nameOut, int nameCap) {
    OObject* o = ((OWorld*)wp)->creationOrder.at(objIdx);
    info6[0] = o->isRobot;
    info6[1] = o->isStatic;
    info6[2] = (int)o->numDofs;
    info6[3] = (int)o->linksCreation.size();
    info6[4] = (int)o->sortedJoints.size();
    info6[5] = o->linksCreation.front()->body->index;
    if (nameOut) { strncpy(nameOut, o->name.c_str(), nameCap - 1); nameOut[nameCap - 1] = 0; }
}

// 16 floats per body: xf.p(2) xf.q.s xf.q.c c(2) a c0(2) a0 v(2) w sleepTime awake type
// Box2DLink::getLocalBoundingBox (:258-261) and Box2DObject::getLocalAABB (:1928-1931, built at :1304-1312): min x, min y, max x, max y
void b2o_world_get_link_aabb(void* wp, int bodyIndex, float* out) {
    OWorld* w = (OWorld*)wp;
    OLink* l = (OLink*)w->b2->bodies.at(bodyIndex)->userData;
    for (int i = 0; i < 4; ++i) out[i] = 0.0f;
    if (!l) return;
    out[0] = l->localAabbMin[0]; out[1] = l->localAabbMin[1]; out[2] = l->localAabbMax[0]; out[3] = l->localAabbMax[1];
}

void b2o_world_get_object_aabb(void* wp, int object, float* out) {
    OWorld* w = (OWorld*)wp;
    OObject* o = w->creationOrder.at(object);
    for (int i = 0; i < 4; ++i) out[i] = o->localAabb[i];
}

// Box2DLink::getPose / getVelocityVector / getCenterOfMass per link: out[n_bodies][8], the ground's row stays 0
void b2o_world_get_link_states(void* wp, float* out) {
    OWorld* w = (OWorld*)wp;
    for (Body* b : w->b2->bodies) {
        OLink* l = (OLink*)b->userData;
        float* o = out + 8 * b->index;
        for (int i = 0; i < 8; ++i) o[i] = 0.0f;
        if (!l) continue;
        l->getPose(o);
        l->getVelocityVector(o + 3);
        l->getCenterOfMass(o + 6);
    }
}

void b2o_world_get_bodies(void* wp, float* out) {
    OWorld* w = (OWorld*)wp;
    for (Body* b : w->b2->bodies) {
        float* o = out + 16 * b->index;
        o[0] = b->xf.p.x; o[1] = b->xf.p.y; o[2] = b->xf.q.s; o[3] = b->xf.q.c;
        o[4] = b->sweep.c.x; o[5] = b->sweep.c.y; o[6] = b->sweep.a;
        o[7] = b->sweep.c0.x; o[8] = b->sweep.c0.y; o[9] = b->sweep.a0;
        o[10] = b->linearVelocity.x; o[11] = b->linearVelocity.y; o[12] = b->angularVelocity;
        o[13] = b->sleepTime; o[14] = b->awake ? 1.0f : 0.0f; o[15] = (float)b->type;
    }
}

// 8 floats per body: mass invMass I invI localCenter(2) type nFixtures
void b2o_world_get_body_model(void* wp, float* out) {
    OWorld* w = (OWorld*)wp;
    for (Body* b : w->b2->bodies) {
        float* o = out + 8 * b->index;
        o[0] = b->mass; o[1] = b->invMass; o[2] = b->I; o[3] = b->invI;
        o[4] = b->sweep.localCenter.x; o[5] = b->sweep.localCenter.y;
        o[6] = (float)b->type; o[7] = (float)b->fixtures.size();
    }
}

// per proxy: ints [body, count, sensor], floats [verts(16) normals(16) centroid(2) friction restitution density radius]
void b2o_world_get_fixture_model(void* wp, int* iout, float* fout) {
    OWorld* w = (OWorld*)wp;
    for (Fixture* f : w->b2->proxies) {
        int* io = iout + 3 * f->proxyId;
        float* fo = fout + 38 * f->proxyId;
        io[0] = f->body->index; io[1] = f->shape.count; io[2] = f->isSensor;
        for (int i = 0; i < 8; ++i) {
            fo[2 * i] = f->shape.vertices[i].x; fo[2 * i + 1] = f->shape.vertices[i].y;
            fo[16 + 2 * i] = f->shape.normals[i].x; fo[16 + 2 * i + 1] = f->shape.normals[i].y;
        }
        fo[32] = f->shape.centroid.x; fo[33] = f->shape.centroid.y;
        fo[34] = f->friction; fo[35] = f->restitution; fo[36] = f->density; fo[37] = f->shape.radius;
    }
}

void b2o_world_get_fat_aabbs(void* wp, float* out) {
    OWorld* w = (OWorld*)wp;
    for (Fixture* f : w->b2->proxies) {
        float* o = out + 4 * f->proxyId;
        o[0] = f->fatAABB.lowerBound.x; o[1] = f->fatAABB.lowerBound.y;
        o[2] = f->fatAABB.upperBound.x; o[3] = f->fatAABB.upperBound.y;
    }
}

// joints in creation order. ints [kind, bodyA, bodyB, limitState, enableMotor, enableLimit];
// floats [linearImpulse(2)/impulse.xy, angularImpulse/impulse.z, motorImpulse, maxMotorTorque, motorSpeed,
//         maxForce, maxTorque, localAnchorA(2), localAnchorB(2), referenceAngle, lower, upper]
void b2o_world_get_joints(void* wp, int* iout, float* fout) {
    OWorld* w = (OWorld*)wp;
    for (Joint* j : w->b2->joints) {
        int* io = iout + 6 * j->index;
        float* fo = fout + 15 * j->index;
        io[0] = j->kind; io[1] = j->bodyA->index; io[2] = j->bodyB->index;
        io[3] = j->limitState; io[4] = j->enableMotor; io[5] = j->enableLimit;
        if (j->kind == kFrictionJoint) {
            fo[0] = j->linearImpulse.x; fo[1] = j->linearImpulse.y; fo[2] = j->angularImpulse; fo[3] = 0.0f;
        } else {
            fo[0] = j->impulse.x; fo[1] = j->impulse.y; fo[2] = j->impulse.z; fo[3] = j->motorImpulse;
        }
        fo[4] = j->maxMotorTorque; fo[5] = j->motorSpeed; fo[6] = j->maxForce; fo[7] = j->maxTorque;
        fo[8] = j->localAnchorA.x; fo[9] = j->localAnchorA.y; fo[10] = j->localAnchorB.x; fo[11] = j->localAnchorB.y;
        fo[12] = j->referenceAngle; fo[13] = j->lowerAngle; fo[14] = j->upperAngle;
        if (j->kind == kPrismaticJoint) {  // motor force and translation limits in the revolute joint's slots
            fo[4] = j->maxMotorForce; fo[13] = j->lowerTranslation; fo[14] = j->upperTranslation;
        }
    }
}

// contacts in creation order (upstream list order is the reverse).
// ints per contact [fixtureA, fixtureB, touching, enabled, pointCount, type, id0, id1];
// floats per contact [localNormal(2) localPoint(2) p0(2) n0 t0 p1(2) n1 t1]
int b2o_world_get_contacts(void* wp, int* iout, float* fout, int maxContacts) {
    OWorld* w = (OWorld*)wp;
    int n = 0;
    for (Contact* c : w->b2->contacts) {
        if (n >= maxContacts) break;
        int* io = iout + 8 * n;
        float* fo = fout + 12 * n;
        io[0] = c->fixtureA->proxyId; io[1] = c->fixtureB->proxyId;
        io[2] = c->touching; io[3] = c->enabled; io[4] = c->manifold.pointCount; io[5] = c->manifold.type;
        io[6] = (int)c->manifold.points[0].id.key; io[7] = (int)c->manifold.points[1].id.key;
        fo[0] = c->manifold.localNormal.x; fo[1] = c->manifold.localNormal.y;
        fo[2] = c->manifold.localPoint.x; fo[3] = c->manifold.localPoint.y;
        for (int k = 0; k < 2; ++k) {
            fo[4 + 4 * k] = c->manifold.points[k].localPoint.x;
            fo[5 + 4 * k] = c->manifold.points[k].localPoint.y;
            fo[6 + 4 * k] = c->manifold.points[k].normalImpulse;
            fo[7 + 4 * k] = c->manifold.points[k].tangentImpulse;
        }
        ++n;
    }
    return (int)w->b2->contacts.size();
}

static int writeContacts(const std::vector<OContact>& cs, int* iout, float* fout, int maxContacts) {
    int n = 0;
    for (const OContact& c : cs) {
        if (n >= maxContacts) break;
        iout[4 * n] = c.bodyA; iout[4 * n + 1] = c.bodyB; iout[4 * n + 2] = c.fixtureA; iout[4 * n + 3] = c.fixtureB;
        for (int k = 0; k < 3; ++k) { fout[6 * n + k] = c.point[k]; fout[6 * n + 3 + k] = c.normal[k]; }
        ++n;
    }
    return (int)cs.size();
}

// checkCollision: touching contacts in upstream list order; ints [bodyA, bodyB, fixtureA, fixtureB], floats [point(3) normal(3)]
int b2o_world_check_collision(void* wp, int* iout, float* fout, int maxContacts) {
    std::vector<OContact> cs;
    ((OWorld*)wp)->updateContactCache(cs);
    return writeContacts(cs, iout, fout, maxContacts);
}

// getObjects(aabb): hits[i] for creation-order object i; isPhysicallyFeasible
void b2o_world_objects_in_aabb(void* wp, const float* aabb, int excludeRobots, int* hits) {
    std::vector<int> h;
    ((OWorld*)wp)->getObjectsInAABB(aabb, excludeRobots != 0, h);
    for (size_t i = 0; i < h.size(); ++i) hits[i] = h[i];
}
int b2o_world_is_physically_feasible(void* wp) { return ((OWorld*)wp)->isPhysicallyFeasible() ? 1 : 0; }

int b2o_world_step_record(void* wp, int steps, int execCtrl, int allowSleep, int* iout, float* fout, int maxContacts) {
    std::vector<OContact> cs;
    ((OWorld*)wp)->stepPhysicsRecord(cs, steps, execCtrl != 0, allowSleep != 0);
    return writeContacts(cs, iout, fout, maxContacts);
}

void b2o_world_set_link_enabled(void* wp, int bodyIndex, int enabled) {
    OWorld* w = (OWorld*)wp;
    OLink* l = (OLink*)w->b2->bodies.at(bodyIndex)->userData;
    if (l) l->setEnabled(enabled != 0);
}

// Box2DLink dynamics parameters (Box2DWorld.cpp:471-566): 0 mass, 1 ground friction coefficient, 2 ground friction torque
// integral, 3 contact friction (settable); 4 getInertia, 5 getCOMInertia, 6 ground friction limit force, 7 limit torque
void b2o_world_set_link_property(void* wp, int bodyIndex, int prop, float value) {
    OWorld* w = (OWorld*)wp;
    OLink* l = (OLink*)w->b2->bodies.at(bodyIndex)->userData;
    if (!l) return;
    switch (prop) {
        case 0: l->setMass(value); break;
        case 1: l->setGroundFrictionCoefficient(value); break;
        case 2: l->setGroundFrictionTorqueIntegral(value); break;
        case 3: l->setContactFriction(value); break;
        default: break;
    }
}

float b2o_world_get_link_property(void* wp, int bodyIndex, int prop) {
    OWorld* w = (OWorld*)wp;
    OLink* l = (OLink*)w->b2->bodies.at(bodyIndex)->userData;
    if (!l) return 0.0f;
    switch (prop) {
        case 0: return l->getMass();
        case 1: return l->groundFriction;
        case 2: return l->groundTorqueIntegral;
        case 3: return l->contactFriction;
        case 4: return l->getInertia();
        case 5: return l->getCOMInertia();
        case 6: return l->getGroundFrictionLimitForce();
        case 7: return l->getGroundFrictionLimitTorque();
        default: return 0.0f;
    }
}

void b2o_world_trace_islands(void* wp, int flag) { ((OWorld*)wp)->b2->traceIslands = flag != 0; }

// flattened island trace of the last step: [nIslands, then per island: nb, bodies..., nj, joints..., nc, (fa,fb)...]
int b2o_world_get_island_trace(void* wp, int* out, int cap) {
    World* b2 = ((OWorld*)wp)->b2.get();
    std::vector<int> v;
    v.push_back((int)b2->traceBodies.size());
    for (size_t i = 0; i < b2->traceBodies.size(); ++i) {
        v.push_back((int)b2->traceBodies[i].size());
        for (int x : b2->traceBodies[i]) v.push_back(x);
        v.push_back((int)b2->traceJoints[i].size());
        for (int x : b2->traceJoints[i]) v.push_back(x);
        v.push_back((int)b2->traceContacts[i].size());
        for (auto& p : b2->traceContacts[i]) { v.push_back(p.first); v.push_back(p.second); }
    }
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// stats of last step: [islands, toiEvents, toiCalls]
void b2o_world_stats(void* wp, int* out3) {
    World* b2 = ((OWorld*)wp)->b2.get();
    out3[0] = b2->stats.islands; out3[1] = b2->stats.toiEvents; out3[2] = b2->stats.toiCalls;
}

// controller introspection: efforts the velocity controller would command now (active DOFs)
int b2o_world_controller_efforts(void* wp, int objIdx, float* out) {
    OWorld* w = (OWorld*)wp;
    OObject* o = w->creationOrder.at(objIdx);
    std::vector<float> vel = o->getDOFVelocities({});
    std::vector<float> eff;
    if (!o->controllerCompute(vel, w->timeStep, eff)) return 0;
    for (size_t i = 0; i < eff.size(); ++i) out[i] = eff[i];
    return (int)eff.size();
}

// ---- trig / math probes for parity of the shared deterministic sincos ------------------------------
void b2o_sincos(const float* x, float* s, float* c, int n) {
    for (int i = 0; i < n; ++i) SinCos(x[i], s[i], c[i]);
}

// ---- polygon probes (Appendix B known answers) -------------------------------------------------------
// in: n points; out: count, vertices(16), normals(16), centroid(2), then mass data for density: mass, cx, cy, I
int b2o_polygon_set(const float* xy, int n, float density, float* out38) {
    std::vector<Vec2> pts;
    for (int i = 0; i < n; ++i) pts.emplace_back(xy[2 * i], xy[2 * i + 1]);
    PolygonShape s;
    s.Set(pts.data(), n);
    for (int i = 0; i < 8; ++i) {
        out38[2 * i] = s.vertices[i].x; out38[2 * i + 1] = s.vertices[i].y;
        out38[16 + 2 * i] = s.normals[i].x; out38[16 + 2 * i + 1] = s.normals[i].y;
    }
    out38[32] = s.centroid.x; out38[33] = s.centroid.y;
    MassData md;
    s.ComputeMass(&md, density);
    out38[34] = md.mass; out38[35] = md.center.x; out38[36] = md.center.y; out38[37] = md.I;
    return s.count;
}

// b2Distance probe: two polygons + transforms (x, y, angle) -> out4 = distance, iterations, |pointA - pointB| check, useRadii
void b2o_distance(const float* xyA, int nA, const float* poseA, const float* xyB, int nB, const float* poseB, int useRadii,
                  float* out6) {
    std::vector<Vec2> a, b;
    for (int i = 0; i < nA; ++i) a.emplace_back(xyA[2 * i], xyA[2 * i + 1]);
    for (int i = 0; i < nB; ++i) b.emplace_back(xyB[2 * i], xyB[2 * i + 1]);
    PolygonShape sa, sb;
    sa.Set(a.data(), nA);
    sb.Set(b.data(), nB);
    DistanceInput in;
    in.proxyA = &sa;
    in.proxyB = &sb;
    in.transformA.Set(Vec2(poseA[0], poseA[1]), poseA[2]);
    in.transformB.Set(Vec2(poseB[0], poseB[1]), poseB[2]);
    in.useRadii = useRadii != 0;
    SimplexCache cache;
    DistanceOutput out;
    DistanceGJK(&out, &cache, &in);
    out6[0] = out.distance; out6[1] = (float)out.iterations;
    out6[2] = out.pointA.x; out6[3] = out.pointA.y; out6[4] = out.pointB.x; out6[5] = out.pointB.y;
}

// b2TimeOfImpact probe: polygon A sweeps from poseA0 to poseA1 (local centre at the origin), B likewise, tMax = 1
// out2 = state (b2TOIOutput::State), t
void b2o_time_of_impact(const float* xyA, int nA, const float* poseA0, const float* poseA1, const float* xyB, int nB,
                        const float* poseB0, const float* poseB1, float* out2) {
    std::vector<Vec2> a, b;
    for (int i = 0; i < nA; ++i) a.emplace_back(xyA[2 * i], xyA[2 * i + 1]);
    for (int i = 0; i < nB; ++i) b.emplace_back(xyB[2 * i], xyB[2 * i + 1]);
    PolygonShape sa, sb;
    sa.Set(a.data(), nA);
    sb.Set(b.data(), nB);
    TOIInput in;
    in.proxyA = &sa;
    in.proxyB = &sb;
    in.sweepA.localCenter = Vec2(0.0f, 0.0f);
    in.sweepA.c0 = Vec2(poseA0[0], poseA0[1]); in.sweepA.a0 = poseA0[2];
    in.sweepA.c = Vec2(poseA1[0], poseA1[1]); in.sweepA.a = poseA1[2];
    in.sweepA.alpha0 = 0.0f;
    in.sweepB.localCenter = Vec2(0.0f, 0.0f);
    in.sweepB.c0 = Vec2(poseB0[0], poseB0[1]); in.sweepB.a0 = poseB0[2];
    in.sweepB.c = Vec2(poseB1[0], poseB1[1]); in.sweepB.a = poseB1[2];
    in.sweepB.alpha0 = 0.0f;
    in.tMax = 1.0f;
    TOIOutput out;
    TimeOfImpact(&out, &in);
    out2[0] = (float)out.state;
    out2[1] = out.t;
}

// narrowphase probe: two polygons + transforms (x, y, angle) -> manifold
// out ints [pointCount, type, id0, id1], floats [localNormal(2) localPoint(2) p0(2) p1(2)]
void b2o_collide_polygons(const float* xyA, int nA, const float* poseA, const float* xyB, int nB, const float* poseB,
                          int* iout, float* fout) {
    std::vector<Vec2> a, b;
    for (int i = 0; i < nA; ++i) a.emplace_back(xyA[2 * i], xyA[2 * i + 1]);
    for (int i = 0; i < nB; ++i) b.emplace_back(xyB[2 * i], xyB[2 * i + 1]);
    PolygonShape sa, sb;
    sa.Set(a.data(), nA);
    sb.Set(b.data(), nB);
    Transform xa, xb;
    xa.Set(Vec2(poseA[0], poseA[1]), poseA[2]);
    xb.Set(Vec2(poseB[0], poseB[1]), poseB[2]);
    Manifold m;
    CollidePolygons(&m, &sa, xa, &sb, xb);
    iout[0] = m.pointCount; iout[1] = m.type;
    iout[2] = (int)m.points[0].id.key; iout[3] = (int)m.points[1].id.key;
    fout[0] = m.localNormal.x; fout[1] = m.localNormal.y; fout[2] = m.localPoint.x; fout[3] = m.localPoint.y;
    fout[4] = m.points[0].localPoint.x; fout[5] = m.points[0].localPoint.y;
    fout[6] = m.points[1].localPoint.x; fout[7] = m.points[1].localPoint.y;
}

// ---- batch driver: one world per thread-pool slot (CPU baseline) ---------------------------
struct OBatch {
    std::vector<OWorld*> worlds;
};

void* b2o_batch_new(void* env, int n) {
    try {
        EnvDesc e = *(EnvDesc*)env;
        finalizeEnv(e);
        OBatch* b = new OBatch();
        for (int i = 0; i < n; ++i) b->worlds.push_back(new OWorld(e));
        return b;
    } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; }
}
void b2o_batch_free(void* bp) {
    OBatch* b = (OBatch*)bp;
    for (OWorld* w : b->worlds) delete w;
    delete b;
}
void* b2o_batch_world(void* bp, int i) { return ((OBatch*)bp)->worlds.at(i); }

// q0/qd0: [n, nq]; targets: [nTargets, n, ndof] for robot object `robotIdx` (or null); each target is
// held for stepsPerTarget steps.  Worlds are dealt round-robin to nThreads std::threads.
int b2o_batch_rollout(void* bp, const float* q0, const float* qd0, int robotIdx, const float* targets, int ndof,
                      int nTargets, int stepsPerTarget, int nThreads, float* qOut, float* qdOut) {
    OBatch* b = (OBatch*)bp;
    int n = (int)b->worlds.size();
    if (n == 0) return 0;
    int nq = b->worlds[0]->stateDim();
    std::atomic<int> failed(0);
    auto work = [&](int tid) {
        for (int i = tid; i < n; i += nThreads) {
            try {
                OWorld* w = b->worlds[i];
                if (q0) w->setState(std::vector<float>(q0 + (size_t)i * nq, q0 + (size_t)(i + 1) * nq),
                                    std::vector<float>(qd0 + (size_t)i * nq, qd0 + (size_t)(i + 1) * nq));
                for (int t = 0; t < nTargets; ++t) {
                    if (targets) {
                        const float* tv = targets + ((size_t)t * n + i) * ndof;
                        w->creationOrder.at(robotIdx)->setTargetVelocity(std::vector<float>(tv, tv + ndof));
                    }
                    w->stepPhysics(stepsPerTarget, true, true);
                }
                if (qOut) {
                    std::vector<float> q, qd;
                    w->getState(q, qd);
                    memcpy(qOut + (size_t)i * nq, q.data(), nq * sizeof(float));
                    memcpy(qdOut + (size_t)i * nq, qd.data(), nq * sizeof(float));
                }
            } catch (...) { failed++; }
        }
    };
    std::vector<std::thread> threads;
    for (int t = 0; t < nThreads; ++t) threads.emplace_back(work, t);
    for (auto& t : threads) t.join();
    return failed.load();
}

}  // extern "C"
