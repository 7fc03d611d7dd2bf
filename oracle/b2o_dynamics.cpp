// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// Box2D 2.3.1 dynamics pipeline restated: b2World::Step =
// [FindNewContacts] -> Collide -> Solve (islands, sequential impulses) -> SolveTOI.
#include "b2o_dynamics.h"

#include <algorithm>

namespace b2o {

// =================================================================================================
// Body / Fixture
// =================================================================================================
World::~World() {
    for (Contact* c : contacts) delete c;
    for (Joint* j : joints) delete j;
    for (Body* b : bodies) {
        for (Fixture* f : b->fixtures) delete f;
        delete b;
    }
}

Body* World::CreateBody(BodyType type, const Vec2& position, float angle) {
    Body* b = new Body();
    b->world = this;
    b->type = type;
    b->index = (int)bodies.size();
    b->xf.p = position;
    b->xf.q.Set(angle);
    b->sweep.localCenter.SetZero();
    b->sweep.c0 = b->xf.p;
    b->sweep.c = b->xf.p;
    b->sweep.a0 = angle;
    b->sweep.a = angle;
    b->sweep.alpha0 = 0.0f;
    if (type == kDynamicBody) { b->mass = 1.0f; b->invMass = 1.0f; }
    bodies.push_back(b);
    return b;
}

Fixture* Body::CreateFixture(const PolygonShape& shape, float density, float friction, float restitution, bool sensor) {
    Fixture* f = new Fixture();
    f->body = this;
    f->shape = shape;
    f->density = density;
    f->friction = friction;
    f->restitution = restitution;
    f->isSensor = sensor;
    // CreateProxies: tight AABB fattened by aabbExtension, queued in the move buffer
    AABB aabb;
    f->shape.ComputeAABB(&aabb, xf);
    Vec2 r(kAabbExtension, kAabbExtension);
    f->fatAABB.lowerBound = aabb.lowerBound - r;
    f->fatAABB.upperBound = aabb.upperBound + r;
    f->proxyId = (int)world->proxies.size();
    world->proxies.push_back(f);
    world->moveBuffer.push_back(f->proxyId);
    fixtures.push_back(f);
    if (f->density > 0.0f) ResetMassData();
    world->newFixture = true;
    return f;
}

void Body::ResetMassData() {
    mass = 0.0f; invMass = 0.0f; I = 0.0f; invI = 0.0f;
    sweep.localCenter.SetZero();
    if (type == kStaticBody || type == kKinematicBody) {
        sweep.c0 = xf.p;
        sweep.c = xf.p;
        sweep.a0 = sweep.a;
        return;
    }
    Vec2 localCenter(0.0f, 0.0f);
    for (int i = (int)fixtures.size() - 1; i >= 0; --i) {  // upstream list: newest first
        Fixture* f = fixtures[i];
        if (f->density == 0.0f) continue;
        MassData md;
        f->shape.ComputeMass(&md, f->density);
        mass += md.mass;
        localCenter += md.mass * md.center;
        I += md.I;
    }
    if (mass > 0.0f) {
        invMass = 1.0f / mass;
        localCenter *= invMass;
    } else {
        mass = 1.0f;
        invMass = 1.0f;
    }
    if (I > 0.0f) {
        I -= mass * Dot(localCenter, localCenter);
        invI = 1.0f / I;
    } else {
        I = 0.0f;
        invI = 0.0f;
    }
    Vec2 oldCenter = sweep.c;
    sweep.localCenter = localCenter;
    sweep.c0 = sweep.c = Mul(xf, sweep.localCenter);
    linearVelocity += Cross(angularVelocity, sweep.c - oldCenter);
}

void Body::GetMassData(MassData* md) const {
    md->mass = mass;
    md->I = I + mass * Dot(sweep.localCenter, sweep.localCenter);
    md->center = sweep.localCenter;
}

void Body::SetMassData(const MassData* md) {
    if (type != kDynamicBody) return;
    invMass = 0.0f; I = 0.0f; invI = 0.0f;
    mass = md->mass;
    if (mass <= 0.0f) mass = 1.0f;
    invMass = 1.0f / mass;
    if (md->I > 0.0f) {
        I = md->I - mass * Dot(md->center, md->center);
        invI = 1.0f / I;
    }
    Vec2 oldCenter = sweep.c;
    sweep.localCenter = md->center;
    sweep.c0 = sweep.c = Mul(xf, sweep.localCenter);
    linearVelocity += Cross(angularVelocity, sweep.c - oldCenter);
}

void World::MoveProxy(Fixture* f, const AABB& aabb, const Vec2& displacement) {
    if (f->fatAABB.Contains(aabb)) return;
    AABB b = aabb;
    Vec2 r(kAabbExtension, kAabbExtension);
    b.lowerBound = b.lowerBound - r;
    b.upperBound = b.upperBound + r;
    Vec2 d = kAabbMultiplier * displacement;
    if (d.x < 0.0f) b.lowerBound.x += d.x; else b.upperBound.x += d.x;
    if (d.y < 0.0f) b.lowerBound.y += d.y; else b.upperBound.y += d.y;
    f->fatAABB = b;
    moveBuffer.push_back(f->proxyId);
}

static void SynchronizeFixture(World* w, Fixture* f, const Transform& xf1, const Transform& xf2) {
    AABB aabb1, aabb2, aabb;
    f->shape.ComputeAABB(&aabb1, xf1);
    f->shape.ComputeAABB(&aabb2, xf2);
    aabb.Combine(aabb1, aabb2);
    Vec2 displacement = xf2.p - xf1.p;
    w->MoveProxy(f, aabb, displacement);
}

void Body::SetTransform(const Vec2& position, float angle) {
    xf.q.Set(angle);
    xf.p = position;
    sweep.c = Mul(xf, sweep.localCenter);
    sweep.a = angle;
    sweep.c0 = sweep.c;
    sweep.a0 = angle;
    for (int i = (int)fixtures.size() - 1; i >= 0; --i) SynchronizeFixture(world, fixtures[i], xf, xf);
    // 2.3.1: new contacts are found at the next FindNewContacts (no call here; 2.3.0 differs)
}

void Body::SynchronizeFixtures() {
    Transform xf1;
    xf1.q.Set(sweep.a0);
    xf1.p = sweep.c0 - Mul(xf1.q, sweep.localCenter);
    for (int i = (int)fixtures.size() - 1; i >= 0; --i) SynchronizeFixture(world, fixtures[i], xf1, xf);
}

void Body::SynchronizeTransform() {
    xf.q.Set(sweep.a);
    xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
}

void Body::Advance(float alpha) {
    sweep.Advance(alpha);
    sweep.c = sweep.c0;
    sweep.a = sweep.a0;
    xf.q.Set(sweep.a);
    xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
}

void Body::SetAwake(bool flag) {
    if (flag) {
        if (!awake) {
            awake = true;
            sleepTime = 0.0f;
        }
    } else {
        awake = false;
        sleepTime = 0.0f;
        linearVelocity.SetZero();
        angularVelocity = 0.0f;
        force.SetZero();
        torque = 0.0f;
    }
}

bool Body::ShouldCollide(const Body* other) const {
    if (type != kDynamicBody && other->type != kDynamicBody) return false;
    for (int i = (int)jointEdges.size() - 1; i >= 0; --i) {
        if (jointEdges[i].other == other) {
            if (jointEdges[i].joint->collideConnected == false) return false;
        }
    }
    return true;
}

void Body::SetLinearVelocity(const Vec2& v) {
    if (type == kStaticBody) return;
    if (Dot(v, v) > 0.0f) SetAwake(true);
    linearVelocity = v;
}
void Body::SetAngularVelocity(float w) {
    if (type == kStaticBody) return;
    if (w * w > 0.0f) SetAwake(true);
    angularVelocity = w;
}
void Body::ApplyForceToCenter(const Vec2& f, bool wake) {
    if (type != kDynamicBody) return;
    if (wake && !awake) SetAwake(true);
    if (awake) force += f;
}
void Body::ApplyTorque(float t, bool wake) {
    if (type != kDynamicBody) return;
    if (wake && !awake) SetAwake(true);
    if (awake) torque += t;
}

// =================================================================================================
// Joints
// =================================================================================================
static void LinkJoint(World* w, Joint* j) {
    j->index = (int)w->joints.size();
    w->joints.push_back(j);
    j->bodyA->jointEdges.push_back(JointEdge{j, j->bodyB});
    j->bodyB->jointEdges.push_back(JointEdge{j, j->bodyA});
    // contacts between the two bodies would be flagged for filtering here; none exist at setup
    if (!j->collideConnected) {
        for (Contact* c : j->bodyB->contacts) {
            Body* oa = c->fixtureA->body;
            Body* ob = c->fixtureB->body;
            if ((oa == j->bodyA && ob == j->bodyB) || (oa == j->bodyB && ob == j->bodyA)) c->filterFlag = true;
        }
    }
}

Joint* World::CreateFrictionJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, float maxForce,
                                  float maxTorque) {
    Joint* j = new Joint();
    j->kind = kFrictionJoint;
    j->bodyA = a;
    j->bodyB = b;
    j->collideConnected = false;
    j->localAnchorA = anchorA;
    j->localAnchorB = anchorB;
    j->maxForce = maxForce;
    j->maxTorque = maxTorque;
    LinkJoint(this, j);
    return j;
}

Joint* World::CreateRevoluteJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, float refAngle,
                                  float lower, float upper, bool enableLimit, bool enableMotor) {
    Joint* j = new Joint();
    j->kind = kRevoluteJoint;
    j->bodyA = a;
    j->bodyB = b;
    j->collideConnected = false;
    j->localAnchorA = anchorA;
    j->localAnchorB = anchorB;
    j->referenceAngle = refAngle;
    j->lowerAngle = lower;
    j->upperAngle = upper;
    j->enableLimit = enableLimit;
    j->enableMotor = enableMotor;
    j->maxMotorTorque = 0.0f;
    j->motorSpeed = 0.0f;
    j->limitState = kInactiveLimit;
    LinkJoint(this, j);
    return j;
}

// b2PrismaticJoint::b2PrismaticJoint (2.3.1), created by the reference at Box2DWorld.cpp:920-940
Joint* World::CreatePrismaticJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, const Vec2& axisA, float refAngle,
                                   float lower, float upper, bool enableLimit, float maxMotorForce_, bool enableMotor_) {
    Joint* j = new Joint();
    j->kind = kPrismaticJoint;
    j->bodyA = a;
    j->bodyB = b;
    j->collideConnected = false;
    j->localAnchorA = anchorA;
    j->localAnchorB = anchorB;
    j->localXAxisA = axisA;
    j->localXAxisA.Normalize();
    j->localYAxisA = Cross(1.0f, j->localXAxisA);
    j->referenceAngle = refAngle;
    j->lowerTranslation = lower;
    j->upperTranslation = upper;
    j->maxMotorForce = maxMotorForce_;
    j->motorSpeed = 0.0f;
    j->enableLimit = enableLimit;
    j->enableMotor = enableMotor_;
    j->limitState = kInactiveLimit;
    LinkJoint(this, j);
    return j;
}

// Solver data: positions/velocities live on the island; joints address them through
// body->islandIndex exactly like upstream's b2SolverData.
namespace {
struct Position { Vec2 c; float a; };
struct Velocity { Vec2 v; float w; };
thread_local Position* g_positions = nullptr;
thread_local Velocity* g_velocities = nullptr;
}  // namespace

// ---- b2PrismaticJoint (2.3.1) -----------------------------------------------------------------------------------------
// Linear constraint (point-to-line):   C = dot(perp, d),  Cdot = dot(perp, vB - vA) + s2 wB - s1 wA
// Angular constraint:                  C = aB - aA - referenceAngle
// Motor / limit along the axis:        Cdot = dot(axis, vB - vA) + a2 wB - a1 wA
void Joint::PrismaticInit(const TimeStep& step) {
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    localCenterA = bodyA->sweep.localCenter;
    localCenterB = bodyB->sweep.localCenter;
    invMassA = bodyA->invMass;
    invMassB = bodyB->invMass;
    invIA = bodyA->invI;
    invIB = bodyB->invI;
    Vec2 cA = g_positions[indexA].c;
    float aA = g_positions[indexA].a;
    Vec2 vA = g_velocities[indexA].v;
    float wA = g_velocities[indexA].w;
    Vec2 cB = g_positions[indexB].c;
    float aB = g_positions[indexB].a;
    Vec2 vB = g_velocities[indexB].v;
    float wB = g_velocities[indexB].w;
    Rot qA(aA), qB(aB);
    rA = Mul(qA, localAnchorA - localCenterA);
    rB = Mul(qB, localAnchorB - localCenterB);
    Vec2 d = (cB - cA) + rB - rA;
    float mA = invMassA, mB = invMassB;
    float iA = invIA, iB = invIB;
    {
        axis = Mul(qA, localXAxisA);
        a1 = Cross(d + rA, axis);
        a2 = Cross(rB, axis);
        motorMass = mA + mB + iA * a1 * a1 + iB * a2 * a2;
        if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
    }
    {
        perp = Mul(qA, localYAxisA);
        s1 = Cross(d + rA, perp);
        s2 = Cross(rB, perp);
        float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
        float k12 = iA * s1 + iB * s2;
        float k13 = iA * s1 * a1 + iB * s2 * a2;
        float k22 = iA + iB;
        if (k22 == 0.0f) k22 = 1.0f;
        float k23 = iA * a1 + iB * a2;
        float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
        massMat.ex = Vec3(k11, k12, k13);
        massMat.ey = Vec3(k12, k22, k23);
        massMat.ez = Vec3(k13, k23, k33);
    }
    if (enableLimit) {
        float jointTranslation = Dot(axis, d);
        if (Abs(upperTranslation - lowerTranslation) < 2.0f * kLinearSlop) {
            limitState = kEqualLimits;
        } else if (jointTranslation <= lowerTranslation) {
            if (limitState != kAtLowerLimit) {
                limitState = kAtLowerLimit;
                impulse.z = 0.0f;
            }
        } else if (jointTranslation >= upperTranslation) {
            if (limitState != kAtUpperLimit) {
                limitState = kAtUpperLimit;
                impulse.z = 0.0f;
            }
        } else {
            limitState = kInactiveLimit;
            impulse.z = 0.0f;
        }
    } else {
        limitState = kInactiveLimit;
        impulse.z = 0.0f;
    }
    if (enableMotor == false) motorImpulse = 0.0f;
    if (step.warmStarting) {
        impulse *= step.dtRatio;
        motorImpulse *= step.dtRatio;
        Vec2 P = impulse.x * perp + (motorImpulse + impulse.z) * axis;
        float LA = impulse.x * s1 + impulse.y + (motorImpulse + impulse.z) * a1;
        float LB = impulse.x * s2 + impulse.y + (motorImpulse + impulse.z) * a2;
        vA -= mA * P;
        wA -= iA * LA;
        vB += mB * P;
        wB += iB * LB;
    } else {
        impulse.SetZero();
        motorImpulse = 0.0f;
    }
    g_velocities[indexA].v = vA;
    g_velocities[indexA].w = wA;
    g_velocities[indexB].v = vB;
    g_velocities[indexB].w = wB;
}

void Joint::PrismaticSolveVelocity(const TimeStep& step) {
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    Vec2 vA = g_velocities[indexA].v;
    float wA = g_velocities[indexA].w;
    Vec2 vB = g_velocities[indexB].v;
    float wB = g_velocities[indexB].w;
    float mA = invMassA, mB = invMassB;
    float iA = invIA, iB = invIB;
    if (enableMotor && limitState != kEqualLimits) {
        float Cdot = Dot(axis, vB - vA) + a2 * wB - a1 * wA;
        float imp = motorMass * (motorSpeed - Cdot);
        float oldImpulse = motorImpulse;
        float maxImpulse = step.dt * maxMotorForce;
        motorImpulse = Clamp(motorImpulse + imp, -maxImpulse, maxImpulse);
        imp = motorImpulse - oldImpulse;
        Vec2 P = imp * axis;
        float LA = imp * a1;
        float LB = imp * a2;
        vA -= mA * P;
        wA -= iA * LA;
        vB += mB * P;
        wB += iB * LB;
    }
    Vec2 Cdot1;
    Cdot1.x = Dot(perp, vB - vA) + s2 * wB - s1 * wA;
    Cdot1.y = wB - wA;
    if (enableLimit && limitState != kInactiveLimit) {
        float Cdot2 = Dot(axis, vB - vA) + a2 * wB - a1 * wA;
        Vec3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
        Vec3 f1 = impulse;
        Vec3 df = massMat.Solve33(-Cdot);
        impulse += df;
        if (limitState == kAtLowerLimit) impulse.z = Max(impulse.z, 0.0f);
        else if (limitState == kAtUpperLimit) impulse.z = Min(impulse.z, 0.0f);
        Vec2 b = -Cdot1 - (impulse.z - f1.z) * Vec2(massMat.ez.x, massMat.ez.y);
        Vec2 f2r = massMat.Solve22(b) + Vec2(f1.x, f1.y);
        impulse.x = f2r.x;
        impulse.y = f2r.y;
        df = Vec3(impulse.x - f1.x, impulse.y - f1.y, impulse.z - f1.z);
        Vec2 P = df.x * perp + df.z * axis;
        float LA = df.x * s1 + df.y + df.z * a1;
        float LB = df.x * s2 + df.y + df.z * a2;
        vA -= mA * P;
        wA -= iA * LA;
        vB += mB * P;
        wB += iB * LB;
    } else {
        Vec2 df = massMat.Solve22(-Cdot1);
        impulse.x += df.x;
        impulse.y += df.y;
        Vec2 P = df.x * perp;
        float LA = df.x * s1 + df.y;
        float LB = df.x * s2 + df.y;
        vA -= mA * P;
        wA -= iA * LA;
        vB += mB * P;
        wB += iB * LB;
    }
    g_velocities[indexA].v = vA;
    g_velocities[indexA].w = wA;
    g_velocities[indexB].v = vB;
    g_velocities[indexB].w = wB;
}

bool Joint::PrismaticSolvePosition() {
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    Vec2 cA = g_positions[indexA].c;
    float aA = g_positions[indexA].a;
    Vec2 cB = g_positions[indexB].c;
    float aB = g_positions[indexB].a;
    Rot qA(aA), qB(aB);
    float mA = invMassA, mB = invMassB;
    float iA = invIA, iB = invIB;
    Vec2 rA_ = Mul(qA, localAnchorA - localCenterA);
    Vec2 rB_ = Mul(qB, localAnchorB - localCenterB);
    Vec2 d = cB + rB_ - cA - rA_;
    Vec2 ax = Mul(qA, localXAxisA);
    float a1_ = Cross(d + rA_, ax);
    float a2_ = Cross(rB_, ax);
    Vec2 pp = Mul(qA, localYAxisA);
    float s1_ = Cross(d + rA_, pp);
    float s2_ = Cross(rB_, pp);
    Vec3 imp;
    Vec2 C1;
    C1.x = Dot(pp, d);
    C1.y = aB - aA - referenceAngle;
    float linearError = Abs(C1.x);
    float angularError = Abs(C1.y);
    bool active = false;
    float C2 = 0.0f;
    if (enableLimit) {
        float translation = Dot(ax, d);
        if (Abs(upperTranslation - lowerTranslation) < 2.0f * kLinearSlop) {
            C2 = Clamp(translation, -kMaxLinearCorrection, kMaxLinearCorrection);
            linearError = Max(linearError, Abs(translation));
            active = true;
        } else if (translation <= lowerTranslation) {
            C2 = Clamp(translation - lowerTranslation + kLinearSlop, -kMaxLinearCorrection, 0.0f);
            linearError = Max(linearError, lowerTranslation - translation);
            active = true;
        } else if (translation >= upperTranslation) {
            C2 = Clamp(translation - upperTranslation - kLinearSlop, 0.0f, kMaxLinearCorrection);
            linearError = Max(linearError, translation - upperTranslation);
            active = true;
        }
    }
    if (active) {
        float k11 = mA + mB + iA * s1_ * s1_ + iB * s2_ * s2_;
        float k12 = iA * s1_ + iB * s2_;
        float k13 = iA * s1_ * a1_ + iB * s2_ * a2_;
        float k22 = iA + iB;
        if (k22 == 0.0f) k22 = 1.0f;
        float k23 = iA * a1_ + iB * a2_;
        float k33 = mA + mB + iA * a1_ * a1_ + iB * a2_ * a2_;
        Mat33 K;
        K.ex = Vec3(k11, k12, k13);
        K.ey = Vec3(k12, k22, k23);
        K.ez = Vec3(k13, k23, k33);
        Vec3 C(C1.x, C1.y, C2);
        imp = K.Solve33(-C);
    } else {
        float k11 = mA + mB + iA * s1_ * s1_ + iB * s2_ * s2_;
        float k12 = iA * s1_ + iB * s2_;
        float k22 = iA + iB;
        if (k22 == 0.0f) k22 = 1.0f;
        Mat22 K;
        K.ex = Vec2(k11, k12);
        K.ey = Vec2(k12, k22);
        Vec2 impulse1 = K.Solve(-C1);
        imp.x = impulse1.x;
        imp.y = impulse1.y;
        imp.z = 0.0f;
    }
    Vec2 P = imp.x * pp + imp.z * ax;
    float LA = imp.x * s1_ + imp.y + imp.z * a1_;
    float LB = imp.x * s2_ + imp.y + imp.z * a2_;
    cA -= mA * P;
    aA -= iA * LA;
    cB += mB * P;
    aB += iB * LB;
    g_positions[indexA].c = cA;
    g_positions[indexA].a = aA;
    g_positions[indexB].c = cB;
    g_positions[indexB].a = aB;
    return linearError <= kLinearSlop && angularError <= kAngularSlop;
}

float Joint::GetJointTranslation() const {
    Vec2 pA = bodyA->GetWorldPoint(localAnchorA);
    Vec2 pB = bodyB->GetWorldPoint(localAnchorB);
    Vec2 d = pB - pA;
    Vec2 ax = Mul(bodyA->xf.q, localXAxisA);
    return Dot(d, ax);
}

float Joint::GetPrismaticJointSpeed() const {
    Vec2 rA_ = Mul(bodyA->xf.q, localAnchorA - bodyA->sweep.localCenter);
    Vec2 rB_ = Mul(bodyB->xf.q, localAnchorB - bodyB->sweep.localCenter);
    Vec2 p1 = bodyA->sweep.c + rA_;
    Vec2 p2 = bodyB->sweep.c + rB_;
    Vec2 d = p2 - p1;
    Vec2 ax = Mul(bodyA->xf.q, localXAxisA);
    Vec2 vA = bodyA->linearVelocity, vB = bodyB->linearVelocity;
    float wA = bodyA->angularVelocity, wB = bodyB->angularVelocity;
    return Dot(d, Cross(wA, ax)) + Dot(ax, vB + Cross(wB, rB_) - vA - Cross(wA, rA_));
}

void Joint::InitVelocityConstraints(const TimeStep& step) {
    if (kind == kPrismaticJoint) return PrismaticInit(step);
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    localCenterA = bodyA->sweep.localCenter;
    localCenterB = bodyB->sweep.localCenter;
    invMassA = bodyA->invMass;
    invMassB = bodyB->invMass;
    invIA = bodyA->invI;
    invIB = bodyB->invI;

    float aA = g_positions[indexA].a;
    Vec2 vA = g_velocities[indexA].v;
    float wA = g_velocities[indexA].w;
    float aB = g_positions[indexB].a;
    Vec2 vB = g_velocities[indexB].v;
    float wB = g_velocities[indexB].w;
    Rot qA(aA), qB(aB);

    rA = Mul(qA, localAnchorA - localCenterA);
    rB = Mul(qB, localAnchorB - localCenterB);
    float mA = invMassA, mB = invMassB;
    float iA = invIA, iB = invIB;

    if (kind == kFrictionJoint) {
        Mat22 K;
        K.ex.x = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        K.ex.y = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        K.ey.x = K.ex.y;
        K.ey.y = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        linearMass = K.GetInverse();
        angularMass = iA + iB;
        if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
        if (step.warmStarting) {
            linearImpulse *= step.dtRatio;
            angularImpulse *= step.dtRatio;
            Vec2 P(linearImpulse.x, linearImpulse.y);
            vA -= mA * P;
            wA -= iA * (Cross(rA, P) + angularImpulse);
            vB += mB * P;
            wB += iB * (Cross(rB, P) + angularImpulse);
        } else {
            linearImpulse.SetZero();
            angularImpulse = 0.0f;
        }
    } else {
        bool fixedRotation = (iA + iB == 0.0f);
        massMat.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        massMat.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        massMat.ez.x = -rA.y * iA - rB.y * iB;
        massMat.ex.y = massMat.ey.x;
        massMat.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        massMat.ez.y = rA.x * iA + rB.x * iB;
        massMat.ex.z = massMat.ez.x;
        massMat.ey.z = massMat.ez.y;
        massMat.ez.z = iA + iB;
        motorMass = iA + iB;
        if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
        if (enableMotor == false || fixedRotation) motorImpulse = 0.0f;
        if (enableLimit && fixedRotation == false) {
            float jointAngle = aB - aA - referenceAngle;
            if (Abs(upperAngle - lowerAngle) < 2.0f * kAngularSlop) {
                limitState = kEqualLimits;
            } else if (jointAngle <= lowerAngle) {
                if (limitState != kAtLowerLimit) impulse.z = 0.0f;
                limitState = kAtLowerLimit;
            } else if (jointAngle >= upperAngle) {
                if (limitState != kAtUpperLimit) impulse.z = 0.0f;
                limitState = kAtUpperLimit;
            } else {
                limitState = kInactiveLimit;
                impulse.z = 0.0f;
            }
        } else {
            limitState = kInactiveLimit;
        }
        if (step.warmStarting) {
            impulse *= step.dtRatio;
            motorImpulse *= step.dtRatio;
            Vec2 P(impulse.x, impulse.y);
            vA -= mA * P;
            wA -= iA * (Cross(rA, P) + motorImpulse + impulse.z);
            vB += mB * P;
            wB += iB * (Cross(rB, P) + motorImpulse + impulse.z);
        } else {
            impulse.SetZero();
            motorImpulse = 0.0f;
        }
    }
    g_velocities[indexA].v = vA;
    g_velocities[indexA].w = wA;
    g_velocities[indexB].v = vB;
    g_velocities[indexB].w = wB;
}

void Joint::SolveVelocityConstraints(const TimeStep& step) {
    if (kind == kPrismaticJoint) return PrismaticSolveVelocity(step);
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    Vec2 vA = g_velocities[indexA].v;
    float wA = g_velocities[indexA].w;
    Vec2 vB = g_velocities[indexB].v;
    float wB = g_velocities[indexB].w;
    float mA = invMassA, mB = invMassB;
    float iA = invIA, iB = invIB;

    if (kind == kFrictionJoint) {
        float h = step.dt;
        {
            float Cdot = wB - wA;
            float imp = -angularMass * Cdot;
            float oldImpulse = angularImpulse;
            float maxImpulse = h * maxTorque;
            angularImpulse = Clamp(angularImpulse + imp, -maxImpulse, maxImpulse);
            imp = angularImpulse - oldImpulse;
            wA -= iA * imp;
            wB += iB * imp;
        }
        {
            Vec2 Cdot = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            Vec2 imp = -Mul(linearMass, Cdot);
            Vec2 oldImpulse = linearImpulse;
            linearImpulse += imp;
            float maxImpulse = h * maxForce;
            if (linearImpulse.LengthSquared() > maxImpulse * maxImpulse) {
                linearImpulse.Normalize();
                linearImpulse *= maxImpulse;
            }
            imp = linearImpulse - oldImpulse;
            vA -= mA * imp;
            wA -= iA * Cross(rA, imp);
            vB += mB * imp;
            wB += iB * Cross(rB, imp);
        }
    } else {
        bool fixedRotation = (iA + iB == 0.0f);
        if (enableMotor && limitState != kEqualLimits && fixedRotation == false) {
            float Cdot = wB - wA - motorSpeed;
            float imp = -motorMass * Cdot;
            float oldImpulse = motorImpulse;
            float maxImpulse = step.dt * maxMotorTorque;
            motorImpulse = Clamp(motorImpulse + imp, -maxImpulse, maxImpulse);
            imp = motorImpulse - oldImpulse;
            wA -= iA * imp;
            wB += iB * imp;
        }
        if (enableLimit && limitState != kInactiveLimit && fixedRotation == false) {
            Vec2 Cdot1 = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            float Cdot2 = wB - wA;
            Vec3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
            Vec3 imp = -massMat.Solve33(Cdot);
            if (limitState == kEqualLimits) {
                impulse += imp;
            } else if (limitState == kAtLowerLimit) {
                float newImpulse = impulse.z + imp.z;
                if (newImpulse < 0.0f) {
                    Vec2 rhs = -Cdot1 + impulse.z * Vec2(massMat.ez.x, massMat.ez.y);
                    Vec2 reduced = massMat.Solve22(rhs);
                    imp.x = reduced.x;
                    imp.y = reduced.y;
                    imp.z = -impulse.z;
                    impulse.x += reduced.x;
                    impulse.y += reduced.y;
                    impulse.z = 0.0f;
                } else {
                    impulse += imp;
                }
            } else if (limitState == kAtUpperLimit) {
                float newImpulse = impulse.z + imp.z;
                if (newImpulse > 0.0f) {
                    Vec2 rhs = -Cdot1 + impulse.z * Vec2(massMat.ez.x, massMat.ez.y);
                    Vec2 reduced = massMat.Solve22(rhs);
                    imp.x = reduced.x;
                    imp.y = reduced.y;
                    imp.z = -impulse.z;
                    impulse.x += reduced.x;
                    impulse.y += reduced.y;
                    impulse.z = 0.0f;
                } else {
                    impulse += imp;
                }
            }
            Vec2 P(imp.x, imp.y);
            vA -= mA * P;
            wA -= iA * (Cross(rA, P) + imp.z);
            vB += mB * P;
            wB += iB * (Cross(rB, P) + imp.z);
        } else {
            Vec2 Cdot = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            Vec2 imp = massMat.Solve22(-Cdot);
            impulse.x += imp.x;
            impulse.y += imp.y;
            vA -= mA * imp;
            wA -= iA * Cross(rA, imp);
            vB += mB * imp;
            wB += iB * Cross(rB, imp);
        }
    }
    g_velocities[indexA].v = vA;
    g_velocities[indexA].w = wA;
    g_velocities[indexB].v = vB;
    g_velocities[indexB].w = wB;
}

bool Joint::SolvePositionConstraints() {
    if (kind == kFrictionJoint) return true;
    if (kind == kPrismaticJoint) return PrismaticSolvePosition();
    int indexA = bodyA->islandIndex, indexB = bodyB->islandIndex;
    Vec2 cA = g_positions[indexA].c;
    float aA = g_positions[indexA].a;
    Vec2 cB = g_positions[indexB].c;
    float aB = g_positions[indexB].a;
    Rot qA(aA), qB(aB);
    float angularError = 0.0f;
    float positionError = 0.0f;
    bool fixedRotation = (invIA + invIB == 0.0f);

    if (enableLimit && limitState != kInactiveLimit && fixedRotation == false) {
        float angle = aB - aA - referenceAngle;
        float limitImpulse = 0.0f;
        if (limitState == kEqualLimits) {
            float C = Clamp(angle - lowerAngle, -kMaxAngularCorrection, kMaxAngularCorrection);
            limitImpulse = -motorMass * C;
            angularError = Abs(C);
        } else if (limitState == kAtLowerLimit) {
            float C = angle - lowerAngle;
            angularError = -C;
            C = Clamp(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
            limitImpulse = -motorMass * C;
        } else if (limitState == kAtUpperLimit) {
            float C = angle - upperAngle;
            angularError = C;
            C = Clamp(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
            limitImpulse = -motorMass * C;
        }
        aA -= invIA * limitImpulse;
        aB += invIB * limitImpulse;
    }
    {
        qA.Set(aA);
        qB.Set(aB);
        Vec2 rA_ = Mul(qA, localAnchorA - localCenterA);
        Vec2 rB_ = Mul(qB, localAnchorB - localCenterB);
        Vec2 C = cB + rB_ - cA - rA_;
        positionError = C.Length();
        float mA = invMassA, mB = invMassB;
        float iA = invIA, iB = invIB;
        Mat22 K;
        K.ex.x = mA + mB + iA * rA_.y * rA_.y + iB * rB_.y * rB_.y;
        K.ex.y = -iA * rA_.x * rA_.y - iB * rB_.x * rB_.y;
        K.ey.x = K.ex.y;
        K.ey.y = mA + mB + iA * rA_.x * rA_.x + iB * rB_.x * rB_.x;
        Vec2 imp = -K.Solve(C);
        cA -= mA * imp;
        aA -= iA * Cross(rA_, imp);
        cB += mB * imp;
        aB += iB * Cross(rB_, imp);
    }
    g_positions[indexA].c = cA;
    g_positions[indexA].a = aA;
    g_positions[indexB].c = cB;
    g_positions[indexB].a = aB;
    return positionError <= kLinearSlop && angularError <= kAngularSlop;
}

// =================================================================================================
// Contacts
// =================================================================================================
void Contact::GetWorldManifold(WorldManifold* wm) const {
    const Body* bodyA = fixtureA->body;
    const Body* bodyB = fixtureB->body;
    wm->Initialize(&manifold, bodyA->xf, fixtureA->shape.radius, bodyB->xf, fixtureB->shape.radius);
}

static bool FilterShouldCollide(const Fixture* fixtureA, const Fixture* fixtureB) {
    const Filter& filterA = fixtureA->filter;
    const Filter& filterB = fixtureB->filter;
    if (filterA.groupIndex == filterB.groupIndex && filterA.groupIndex != 0) return filterA.groupIndex > 0;
    bool collide = (filterA.maskBits & filterB.categoryBits) != 0 && (filterA.categoryBits & filterB.maskBits) != 0;
    return collide;
}

void World::AddPair(Fixture* fixtureA, Fixture* fixtureB) {
    Body* bodyA = fixtureA->body;
    Body* bodyB = fixtureB->body;
    if (bodyA == bodyB) return;
    // does a contact already exist? (walk bodyB's contact edges)
    for (int i = (int)bodyB->contacts.size() - 1; i >= 0; --i) {
        Contact* c = bodyB->contacts[i];
        Body* other = c->fixtureA->body == bodyB ? c->fixtureB->body : c->fixtureA->body;
        if (other == bodyA) {
            Fixture* fA = c->fixtureA;
            Fixture* fB = c->fixtureB;
            if (fA == fixtureA && fB == fixtureB) return;
            if (fA == fixtureB && fB == fixtureA) return;
        }
    }
    if (bodyB->ShouldCollide(bodyA) == false) return;
    if (FilterShouldCollide(fixtureA, fixtureB) == false) return;

    Contact* c = new Contact();
    c->fixtureA = fixtureA;
    c->fixtureB = fixtureB;
    c->enabled = true;
    c->manifold.pointCount = 0;
    c->toiCount = 0;
    c->friction = sqrtf(fixtureA->friction * fixtureB->friction);
    c->restitution = fixtureA->restitution > fixtureB->restitution ? fixtureA->restitution : fixtureB->restitution;
    c->serial = contactSerial++;
    contacts.push_back(c);
    bodyA->contacts.push_back(c);
    bodyB->contacts.push_back(c);
    if (fixtureA->isSensor == false && fixtureB->isSensor == false) {
        bodyA->SetAwake(true);
        bodyB->SetAwake(true);
    }
}

void World::FindNewContacts() {
    // b2BroadPhase::UpdatePairs: every moved proxy queries all fat-overlapping proxies; pairs are
    // sorted (proxyIdA, proxyIdB) and de-duplicated before AddPair.
    std::vector<std::pair<int, int>> pairs;
    for (int queryId : moveBuffer) {
        const AABB& fat = proxies[queryId]->fatAABB;
        for (Fixture* f : proxies) {
            if (f->proxyId == queryId) continue;
            if (TestOverlap(f->fatAABB, fat)) {
                pairs.emplace_back(Min(f->proxyId, queryId), Max(f->proxyId, queryId));
            }
        }
    }
    moveBuffer.clear();
    std::sort(pairs.begin(), pairs.end());
    size_t i = 0;
    while (i < pairs.size()) {
        std::pair<int, int> primary = pairs[i];
        AddPair(proxies[primary.first], proxies[primary.second]);
        ++i;
        while (i < pairs.size() && pairs[i] == primary) ++i;
    }
}

void World::DestroyContact(Contact* c) {
    Body* bodyA = c->fixtureA->body;
    Body* bodyB = c->fixtureB->body;
    if (listener && c->IsTouching()) listener->EndContact(c);
    contacts.erase(std::find(contacts.begin(), contacts.end(), c));
    bodyA->contacts.erase(std::find(bodyA->contacts.begin(), bodyA->contacts.end(), c));
    bodyB->contacts.erase(std::find(bodyB->contacts.begin(), bodyB->contacts.end(), c));
    // b2Contact::Destroy
    if (c->manifold.pointCount > 0 && c->fixtureA->isSensor == false && c->fixtureB->isSensor == false) {
        bodyA->SetAwake(true);
        bodyB->SetAwake(true);
    }
    delete c;
}

void World::UpdateContact(Contact* c) {
    Manifold oldManifold = c->manifold;
    c->enabled = true;
    bool touching = false;
    bool wasTouching = c->touching;
    bool sensor = c->fixtureA->isSensor || c->fixtureB->isSensor;
    Body* bodyA = c->fixtureA->body;
    Body* bodyB = c->fixtureB->body;
    const Transform& xfA = bodyA->xf;
    const Transform& xfB = bodyB->xf;
    if (sensor) {
        touching = TestOverlapShapes(&c->fixtureA->shape, &c->fixtureB->shape, xfA, xfB);
        c->manifold.pointCount = 0;
    } else {
        CollidePolygons(&c->manifold, &c->fixtureA->shape, xfA, &c->fixtureB->shape, xfB);
        touching = c->manifold.pointCount > 0;
        for (int i = 0; i < c->manifold.pointCount; ++i) {
            ManifoldPoint* mp2 = c->manifold.points + i;
            mp2->normalImpulse = 0.0f;
            mp2->tangentImpulse = 0.0f;
            ContactID id2 = mp2->id;
            for (int j = 0; j < oldManifold.pointCount; ++j) {
                ManifoldPoint* mp1 = oldManifold.points + j;
                if (mp1->id.key == id2.key) {
                    mp2->normalImpulse = mp1->normalImpulse;
                    mp2->tangentImpulse = mp1->tangentImpulse;
                    break;
                }
            }
        }
        if (touching != wasTouching) {
            bodyA->SetAwake(true);
            bodyB->SetAwake(true);
        }
    }
    c->touching = touching;
    if (wasTouching == false && touching == true && listener) listener->BeginContact(c);
    if (wasTouching == true && touching == false && listener) listener->EndContact(c);
}

void World::Collide() {
    // iterate in upstream list order (newest first); contacts may be destroyed while iterating
    int i = (int)contacts.size() - 1;
    while (i >= 0) {
        Contact* c = contacts[i];
        Fixture* fixtureA = c->fixtureA;
        Fixture* fixtureB = c->fixtureB;
        Body* bodyA = fixtureA->body;
        Body* bodyB = fixtureB->body;
        if (c->filterFlag) {
            if (bodyB->ShouldCollide(bodyA) == false) { DestroyContact(c); --i; continue; }
            if (FilterShouldCollide(fixtureA, fixtureB) == false) { DestroyContact(c); --i; continue; }
            c->filterFlag = false;
        }
        bool activeA = bodyA->IsAwake() && bodyA->type != kStaticBody;
        bool activeB = bodyB->IsAwake() && bodyB->type != kStaticBody;
        if (activeA == false && activeB == false) { --i; continue; }
        bool overlap = TestOverlap(fixtureA->fatAABB, fixtureB->fatAABB);
        if (overlap == false) { DestroyContact(c); --i; continue; }
        UpdateContact(c);
        --i;
    }
}

void World::UpdateContacts() {
    FindNewContacts();
    Collide();
}

// =================================================================================================
// Contact solver (b2ContactSolver)
// =================================================================================================
namespace {
struct VelocityConstraintPoint {
    Vec2 rA, rB;
    float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias;
};
struct ContactVelocityConstraint {
    VelocityConstraintPoint points[kMaxManifoldPoints];
    Vec2 normal;
    Mat22 normalMass;
    Mat22 K;
    int indexA, indexB;
    float invMassA, invMassB, invIA, invIB;
    float friction, restitution, tangentSpeed;
    int pointCount;
    int contactIndex;
};
struct ContactPositionConstraint {
    Vec2 localPoints[kMaxManifoldPoints];
    Vec2 localNormal, localPoint;
    int indexA, indexB;
    float invMassA, invMassB;
    Vec2 localCenterA, localCenterB;
    float invIA, invIB;
    int type;
    float radiusA, radiusB;
    int pointCount;
};

struct ContactSolver {
    TimeStep step;
    Position* positions;
    Velocity* velocities;
    std::vector<ContactVelocityConstraint> vcs;
    std::vector<ContactPositionConstraint> pcs;
    std::vector<Contact*>& contacts;

    ContactSolver(const TimeStep& s, std::vector<Contact*>& cs, Position* p, Velocity* v)
        : step(s), positions(p), velocities(v), contacts(cs) {
        int count = (int)contacts.size();
        vcs.resize(count);
        pcs.resize(count);
        for (int i = 0; i < count; ++i) {
            Contact* contact = contacts[i];
            Fixture* fixtureA = contact->fixtureA;
            Fixture* fixtureB = contact->fixtureB;
            float radiusA = fixtureA->shape.radius;
            float radiusB = fixtureB->shape.radius;
            Body* bodyA = fixtureA->body;
            Body* bodyB = fixtureB->body;
            Manifold* manifold = &contact->manifold;
            int pointCount = manifold->pointCount;

            ContactVelocityConstraint* vc = &vcs[i];
            vc->friction = contact->friction;
            vc->restitution = contact->restitution;
            vc->tangentSpeed = 0.0f;
            vc->indexA = bodyA->islandIndex;
            vc->indexB = bodyB->islandIndex;
            vc->invMassA = bodyA->invMass;
            vc->invMassB = bodyB->invMass;
            vc->invIA = bodyA->invI;
            vc->invIB = bodyB->invI;
            vc->contactIndex = i;
            vc->pointCount = pointCount;
            vc->K.SetZero();
            vc->normalMass.SetZero();

            ContactPositionConstraint* pc = &pcs[i];
            pc->indexA = bodyA->islandIndex;
            pc->indexB = bodyB->islandIndex;
            pc->invMassA = bodyA->invMass;
            pc->invMassB = bodyB->invMass;
            pc->localCenterA = bodyA->sweep.localCenter;
            pc->localCenterB = bodyB->sweep.localCenter;
            pc->invIA = bodyA->invI;
            pc->invIB = bodyB->invI;
            pc->localNormal = manifold->localNormal;
            pc->localPoint = manifold->localPoint;
            pc->pointCount = pointCount;
            pc->radiusA = radiusA;
            pc->radiusB = radiusB;
            pc->type = manifold->type;

            for (int j = 0; j < pointCount; ++j) {
                ManifoldPoint* cp = manifold->points + j;
                VelocityConstraintPoint* vcp = vc->points + j;
                if (step.warmStarting) {
                    vcp->normalImpulse = step.dtRatio * cp->normalImpulse;
                    vcp->tangentImpulse = step.dtRatio * cp->tangentImpulse;
                } else {
                    vcp->normalImpulse = 0.0f;
                    vcp->tangentImpulse = 0.0f;
                }
                vcp->rA.SetZero();
                vcp->rB.SetZero();
                vcp->normalMass = 0.0f;
                vcp->tangentMass = 0.0f;
                vcp->velocityBias = 0.0f;
                pc->localPoints[j] = cp->localPoint;
            }
        }
    }

    void InitializeVelocityConstraints() {
        for (size_t i = 0; i < vcs.size(); ++i) {
            ContactVelocityConstraint* vc = &vcs[i];
            ContactPositionConstraint* pc = &pcs[i];
            float radiusA = pc->radiusA;
            float radiusB = pc->radiusB;
            Manifold* manifold = &contacts[vc->contactIndex]->manifold;
            int indexA = vc->indexA;
            int indexB = vc->indexB;
            float mA = vc->invMassA;
            float mB = vc->invMassB;
            float iA = vc->invIA;
            float iB = vc->invIB;
            Vec2 localCenterA = pc->localCenterA;
            Vec2 localCenterB = pc->localCenterB;
            Vec2 cA = positions[indexA].c;
            float aA = positions[indexA].a;
            Vec2 vA = velocities[indexA].v;
            float wA = velocities[indexA].w;
            Vec2 cB = positions[indexB].c;
            float aB = positions[indexB].a;
            Vec2 vB = velocities[indexB].v;
            float wB = velocities[indexB].w;

            Transform xfA, xfB;
            xfA.q.Set(aA);
            xfB.q.Set(aB);
            xfA.p = cA - Mul(xfA.q, localCenterA);
            xfB.p = cB - Mul(xfB.q, localCenterB);

            WorldManifold worldManifold;
            worldManifold.Initialize(manifold, xfA, radiusA, xfB, radiusB);
            vc->normal = worldManifold.normal;

            int pointCount = vc->pointCount;
            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint* vcp = vc->points + j;
                vcp->rA = worldManifold.points[j] - cA;
                vcp->rB = worldManifold.points[j] - cB;
                float rnA = Cross(vcp->rA, vc->normal);
                float rnB = Cross(vcp->rB, vc->normal);
                float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
                Vec2 tangent = Cross(vc->normal, 1.0f);
                float rtA = Cross(vcp->rA, tangent);
                float rtB = Cross(vcp->rB, tangent);
                float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
                vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
                vcp->velocityBias = 0.0f;
                float vRel = Dot(vc->normal, vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA));
                if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
            }
            if (vc->pointCount == 2) {
                VelocityConstraintPoint* vcp1 = vc->points + 0;
                VelocityConstraintPoint* vcp2 = vc->points + 1;
                float rn1A = Cross(vcp1->rA, vc->normal);
                float rn1B = Cross(vcp1->rB, vc->normal);
                float rn2A = Cross(vcp2->rA, vc->normal);
                float rn2B = Cross(vcp2->rB, vc->normal);
                float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
                float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
                float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
                const float k_maxConditionNumber = 1000.0f;
                if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                    vc->K.ex.Set(k11, k12);
                    vc->K.ey.Set(k12, k22);
                    vc->normalMass = vc->K.GetInverse();
                } else {
                    vc->pointCount = 1;
                }
            }
        }
    }

    void WarmStart() {
        for (size_t i = 0; i < vcs.size(); ++i) {
            ContactVelocityConstraint* vc = &vcs[i];
            int indexA = vc->indexA, indexB = vc->indexB;
            float mA = vc->invMassA, iA = vc->invIA;
            float mB = vc->invMassB, iB = vc->invIB;
            int pointCount = vc->pointCount;
            Vec2 vA = velocities[indexA].v;
            float wA = velocities[indexA].w;
            Vec2 vB = velocities[indexB].v;
            float wB = velocities[indexB].w;
            Vec2 normal = vc->normal;
            Vec2 tangent = Cross(normal, 1.0f);
            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint* vcp = vc->points + j;
                Vec2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
                wA -= iA * Cross(vcp->rA, P);
                vA -= mA * P;
                wB += iB * Cross(vcp->rB, P);
                vB += mB * P;
            }
            velocities[indexA].v = vA;
            velocities[indexA].w = wA;
            velocities[indexB].v = vB;
            velocities[indexB].w = wB;
        }
    }

    void SolveVelocityConstraints() {
        for (size_t i = 0; i < vcs.size(); ++i) {
            ContactVelocityConstraint* vc = &vcs[i];
            int indexA = vc->indexA, indexB = vc->indexB;
            float mA = vc->invMassA, iA = vc->invIA;
            float mB = vc->invMassB, iB = vc->invIB;
            int pointCount = vc->pointCount;
            Vec2 vA = velocities[indexA].v;
            float wA = velocities[indexA].w;
            Vec2 vB = velocities[indexB].v;
            float wB = velocities[indexB].w;
            Vec2 normal = vc->normal;
            Vec2 tangent = Cross(normal, 1.0f);
            float friction = vc->friction;

            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint* vcp = vc->points + j;
                Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
                float vt = Dot(dv, tangent) - vc->tangentSpeed;
                float lambda = vcp->tangentMass * (-vt);
                float maxFriction = friction * vcp->normalImpulse;
                float newImpulse = Clamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - vcp->tangentImpulse;
                vcp->tangentImpulse = newImpulse;
                Vec2 P = lambda * tangent;
                vA -= mA * P;
                wA -= iA * Cross(vcp->rA, P);
                vB += mB * P;
                wB += iB * Cross(vcp->rB, P);
            }

            if (vc->pointCount == 1) {
                VelocityConstraintPoint* vcp = vc->points + 0;
                Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
                float vn = Dot(dv, normal);
                float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
                float newImpulse = Max(vcp->normalImpulse + lambda, 0.0f);
                lambda = newImpulse - vcp->normalImpulse;
                vcp->normalImpulse = newImpulse;
                Vec2 P = lambda * normal;
                vA -= mA * P;
                wA -= iA * Cross(vcp->rA, P);
                vB += mB * P;
                wB += iB * Cross(vcp->rB, P);
            } else {
                VelocityConstraintPoint* cp1 = vc->points + 0;
                VelocityConstraintPoint* cp2 = vc->points + 1;
                Vec2 a(cp1->normalImpulse, cp2->normalImpulse);
                Vec2 dv1 = vB + Cross(wB, cp1->rB) - vA - Cross(wA, cp1->rA);
                Vec2 dv2 = vB + Cross(wB, cp2->rB) - vA - Cross(wA, cp2->rA);
                float vn1 = Dot(dv1, normal);
                float vn2 = Dot(dv2, normal);
                Vec2 b;
                b.x = vn1 - cp1->velocityBias;
                b.y = vn2 - cp2->velocityBias;
                b -= Mul(vc->K, a);
                for (;;) {
                    Vec2 x = -Mul(vc->normalMass, b);
                    if (x.x >= 0.0f && x.y >= 0.0f) {
                        Vec2 d = x - a;
                        Vec2 P1 = d.x * normal;
                        Vec2 P2 = d.y * normal;
                        vA -= mA * (P1 + P2);
                        wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
                        vB += mB * (P1 + P2);
                        wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
                        cp1->normalImpulse = x.x;
                        cp2->normalImpulse = x.y;
                        break;
                    }
                    x.x = -cp1->normalMass * b.x;
                    x.y = 0.0f;
                    vn1 = 0.0f;
                    vn2 = vc->K.ex.y * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) {
                        Vec2 d = x - a;
                        Vec2 P1 = d.x * normal;
                        Vec2 P2 = d.y * normal;
                        vA -= mA * (P1 + P2);
                        wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
                        vB += mB * (P1 + P2);
                        wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
                        cp1->normalImpulse = x.x;
                        cp2->normalImpulse = x.y;
                        break;
                    }
                    x.x = 0.0f;
                    x.y = -cp2->normalMass * b.y;
                    vn1 = vc->K.ey.x * x.y + b.x;
                    vn2 = 0.0f;
                    if (x.y >= 0.0f && vn1 >= 0.0f) {
                        Vec2 d = x - a;
                        Vec2 P1 = d.x * normal;
                        Vec2 P2 = d.y * normal;
                        vA -= mA * (P1 + P2);
                        wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
                        vB += mB * (P1 + P2);
                        wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
                        cp1->normalImpulse = x.x;
                        cp2->normalImpulse = x.y;
                        break;
                    }
                    x.x = 0.0f;
                    x.y = 0.0f;
                    vn1 = b.x;
                    vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) {
                        Vec2 d = x - a;
                        Vec2 P1 = d.x * normal;
                        Vec2 P2 = d.y * normal;
                        vA -= mA * (P1 + P2);
                        wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
                        vB += mB * (P1 + P2);
                        wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
                        cp1->normalImpulse = x.x;
                        cp2->normalImpulse = x.y;
                        break;
                    }
                    break;
                }
            }
            velocities[indexA].v = vA;
            velocities[indexA].w = wA;
            velocities[indexB].v = vB;
            velocities[indexB].w = wB;
        }
    }

    void StoreImpulses() {
        for (size_t i = 0; i < vcs.size(); ++i) {
            ContactVelocityConstraint* vc = &vcs[i];
            Manifold* manifold = &contacts[vc->contactIndex]->manifold;
            for (int j = 0; j < vc->pointCount; ++j) {
                manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
                manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
            }
        }
    }

    static void PositionManifold(const ContactPositionConstraint* pc, const Transform& xfA, const Transform& xfB,
                                 int index, Vec2* normal, Vec2* point, float* separation) {
        switch (pc->type) {
        case Manifold::e_faceA: {
            *normal = Mul(xfA.q, pc->localNormal);
            Vec2 planePoint = Mul(xfA, pc->localPoint);
            Vec2 clipPoint = Mul(xfB, pc->localPoints[index]);
            *separation = Dot(clipPoint - planePoint, *normal) - pc->radiusA - pc->radiusB;
            *point = clipPoint;
        } break;
        case Manifold::e_faceB: {
            *normal = Mul(xfB.q, pc->localNormal);
            Vec2 planePoint = Mul(xfB, pc->localPoint);
            Vec2 clipPoint = Mul(xfA, pc->localPoints[index]);
            *separation = Dot(clipPoint - planePoint, *normal) - pc->radiusA - pc->radiusB;
            *point = clipPoint;
            *normal = -(*normal);
        } break;
        default: break;
        }
    }

    // toiIndexA/B < 0: regular position solve; otherwise the TOI variant
    bool SolvePositions(int toiIndexA, int toiIndexB, bool toi) {
        float minSeparation = 0.0f;
        for (size_t i = 0; i < pcs.size(); ++i) {
            ContactPositionConstraint* pc = &pcs[i];
            int indexA = pc->indexA, indexB = pc->indexB;
            Vec2 localCenterA = pc->localCenterA;
            Vec2 localCenterB = pc->localCenterB;
            int pointCount = pc->pointCount;
            float mA, iA, mB, iB;
            if (toi) {
                mA = 0.0f; iA = 0.0f;
                if (indexA == toiIndexA || indexA == toiIndexB) { mA = pc->invMassA; iA = pc->invIA; }
                mB = 0.0f; iB = 0.0f;
                if (indexB == toiIndexA || indexB == toiIndexB) { mB = pc->invMassB; iB = pc->invIB; }
            } else {
                mA = pc->invMassA; iA = pc->invIA;
                mB = pc->invMassB; iB = pc->invIB;
            }
            Vec2 cA = positions[indexA].c;
            float aA = positions[indexA].a;
            Vec2 cB = positions[indexB].c;
            float aB = positions[indexB].a;
            for (int j = 0; j < pointCount; ++j) {
                Transform xfA, xfB;
                xfA.q.Set(aA);
                xfB.q.Set(aB);
                xfA.p = cA - Mul(xfA.q, localCenterA);
                xfB.p = cB - Mul(xfB.q, localCenterB);
                Vec2 normal, point;
                float separation = 0.0f;
                PositionManifold(pc, xfA, xfB, j, &normal, &point, &separation);
                Vec2 rA = point - cA;
                Vec2 rB = point - cB;
                minSeparation = Min(minSeparation, separation);
                float C = Clamp((toi ? kToiBaumgarte : kBaumgarte) * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                float rnA = Cross(rA, normal);
                float rnB = Cross(rB, normal);
                float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                Vec2 P = impulse * normal;
                cA -= mA * P;
                aA -= iA * Cross(rA, P);
                cB += mB * P;
                aB += iB * Cross(rB, P);
            }
            positions[indexA].c = cA;
            positions[indexA].a = aA;
            positions[indexB].c = cB;
            positions[indexB].a = aB;
        }
        return minSeparation >= (toi ? -1.5f * kLinearSlop : -3.0f * kLinearSlop);
    }
};
}  // namespace

// =================================================================================================
// World stepping
// =================================================================================================
void World::SetAllowSleeping(bool flag) {
    if (flag == allowSleep) return;
    allowSleep = flag;
    if (allowSleep == false) {
        for (int i = (int)bodies.size() - 1; i >= 0; --i) bodies[i]->SetAwake(true);
    }
}

void World::ClearForces() {
    for (Body* b : bodies) {
        b->force.SetZero();
        b->torque = 0.0f;
    }
}

static void IslandSolve(World* world, std::vector<Body*>& ibodies, std::vector<Contact*>& icontacts,
                        std::vector<Joint*>& ijoints, const TimeStep& step, const Vec2& gravity, bool allowSleep) {
    float h = step.dt;
    std::vector<Position> positions(ibodies.size());
    std::vector<Velocity> velocities(ibodies.size());
    for (size_t i = 0; i < ibodies.size(); ++i) {
        Body* b = ibodies[i];
        Vec2 c = b->sweep.c;
        float a = b->sweep.a;
        Vec2 v = b->linearVelocity;
        float w = b->angularVelocity;
        b->sweep.c0 = b->sweep.c;
        b->sweep.a0 = b->sweep.a;
        if (b->type == kDynamicBody) {
            v += h * (1.0f * gravity + b->invMass * b->force);
            w += h * b->invI * b->torque;
            v *= 1.0f / (1.0f + h * 0.0f);
            w *= 1.0f / (1.0f + h * 0.0f);
        }
        positions[i].c = c;
        positions[i].a = a;
        velocities[i].v = v;
        velocities[i].w = w;
    }
    g_positions = positions.data();
    g_velocities = velocities.data();

    ContactSolver contactSolver(step, icontacts, positions.data(), velocities.data());
    contactSolver.InitializeVelocityConstraints();
    if (step.warmStarting) contactSolver.WarmStart();
    for (Joint* j : ijoints) j->InitVelocityConstraints(step);

    for (int i = 0; i < step.velocityIterations; ++i) {
        for (Joint* j : ijoints) j->SolveVelocityConstraints(step);
        contactSolver.SolveVelocityConstraints();
    }
    contactSolver.StoreImpulses();

    for (size_t i = 0; i < ibodies.size(); ++i) {
        Vec2 c = positions[i].c;
        float a = positions[i].a;
        Vec2 v = velocities[i].v;
        float w = velocities[i].w;
        Vec2 translation = h * v;
        if (Dot(translation, translation) > kMaxTranslationSquared) {
            float ratio = kMaxTranslation / translation.Length();
            v *= ratio;
        }
        float rotation = h * w;
        if (rotation * rotation > kMaxRotationSquared) {
            float ratio = kMaxRotation / Abs(rotation);
            w *= ratio;
        }
        c += h * v;
        a += h * w;
        positions[i].c = c;
        positions[i].a = a;
        velocities[i].v = v;
        velocities[i].w = w;
    }

    bool positionSolved = false;
    for (int i = 0; i < step.positionIterations; ++i) {
        bool contactsOkay = contactSolver.SolvePositions(-1, -1, false);
        bool jointsOkay = true;
        for (Joint* j : ijoints) {
            bool jointOkay = j->SolvePositionConstraints();
            jointsOkay = jointsOkay && jointOkay;
        }
        if (contactsOkay && jointsOkay) {
            positionSolved = true;
            break;
        }
    }

    for (size_t i = 0; i < ibodies.size(); ++i) {
        Body* body = ibodies[i];
        body->sweep.c = positions[i].c;
        body->sweep.a = positions[i].a;
        body->linearVelocity = velocities[i].v;
        body->angularVelocity = velocities[i].w;
        body->SynchronizeTransform();
    }

    if (allowSleep) {
        float minSleepTime = kMaxFloat;
        const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
        const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
        for (Body* b : ibodies) {
            if (b->type == kStaticBody) continue;
            if (b->autoSleep == false || b->angularVelocity * b->angularVelocity > angTolSqr ||
                Dot(b->linearVelocity, b->linearVelocity) > linTolSqr) {
                b->sleepTime = 0.0f;
                minSleepTime = 0.0f;
            } else {
                b->sleepTime += h;
                minSleepTime = Min(minSleepTime, b->sleepTime);
            }
        }
        if (minSleepTime >= kTimeToSleep && positionSolved) {
            for (Body* b : ibodies) b->SetAwake(false);
        }
    }
    (void)world;
    g_positions = nullptr;
    g_velocities = nullptr;
}

void World::Solve(const TimeStep& step) {
    for (Body* b : bodies) b->islandFlag = false;
    for (Contact* c : contacts) c->islandFlag = false;
    for (Joint* j : joints) j->islandFlag = false;
    if (traceIslands) { traceBodies.clear(); traceJoints.clear(); traceContacts.clear(); }
    stats.islands = 0;

    std::vector<Body*> stack;
    std::vector<Body*> ibodies;
    std::vector<Contact*> icontacts;
    std::vector<Joint*> ijoints;
    for (int s = (int)bodies.size() - 1; s >= 0; --s) {  // upstream body list: newest first
        Body* seed = bodies[s];
        if (seed->islandFlag) continue;
        if (seed->IsAwake() == false) continue;
        if (seed->type == kStaticBody) continue;

        ibodies.clear();
        icontacts.clear();
        ijoints.clear();
        stack.clear();
        stack.push_back(seed);
        seed->islandFlag = true;

        while (!stack.empty()) {
            Body* b = stack.back();
            stack.pop_back();
            b->islandIndex = (int)ibodies.size();
            ibodies.push_back(b);
            b->SetAwake(true);
            if (b->type == kStaticBody) continue;

            for (int i = (int)b->contacts.size() - 1; i >= 0; --i) {  // contact edges: newest first
                Contact* contact = b->contacts[i];
                if (contact->islandFlag) continue;
                if (contact->IsEnabled() == false || contact->IsTouching() == false) continue;
                if (contact->fixtureA->isSensor || contact->fixtureB->isSensor) continue;
                icontacts.push_back(contact);
                contact->islandFlag = true;
                Body* other = contact->fixtureA->body == b ? contact->fixtureB->body : contact->fixtureA->body;
                if (other->islandFlag) continue;
                stack.push_back(other);
                other->islandFlag = true;
            }
            for (int i = (int)b->jointEdges.size() - 1; i >= 0; --i) {  // joint edges: newest first
                JointEdge& je = b->jointEdges[i];
                if (je.joint->islandFlag == true) continue;
                Body* other = je.other;
                ijoints.push_back(je.joint);
                je.joint->islandFlag = true;
                if (other->islandFlag) continue;
                stack.push_back(other);
                other->islandFlag = true;
            }
        }

        if (traceIslands) {
            std::vector<int> tb, tj;
            std::vector<std::pair<int, int>> tc;
            for (Body* b : ibodies) tb.push_back(b->index);
            for (Joint* j : ijoints) tj.push_back(j->index);
            for (Contact* c : icontacts) tc.emplace_back(c->fixtureA->proxyId, c->fixtureB->proxyId);
            traceBodies.push_back(tb);
            traceJoints.push_back(tj);
            traceContacts.push_back(tc);
        }
        ++stats.islands;
        IslandSolve(this, ibodies, icontacts, ijoints, step, gravity, allowSleep);

        for (Body* b : ibodies) {
            if (b->type == kStaticBody) b->islandFlag = false;
        }
    }

    for (int i = (int)bodies.size() - 1; i >= 0; --i) {
        Body* b = bodies[i];
        if (b->islandFlag == false) continue;
        if (b->type == kStaticBody) continue;
        b->SynchronizeFixtures();
    }
    FindNewContacts();
}

void World::SolveTOI(const TimeStep& step) {
    if (stepComplete) {
        for (Body* b : bodies) {
            b->islandFlag = false;
            b->sweep.alpha0 = 0.0f;
        }
        for (Contact* c : contacts) {
            c->toiFlag = false;
            c->islandFlag = false;
            c->toiCount = 0;
            c->toi = 1.0f;
        }
    }

    for (;;) {
        Contact* minContact = nullptr;
        float minAlpha = 1.0f;
        for (int ci = (int)contacts.size() - 1; ci >= 0; --ci) {  // list order: newest first
            Contact* c = contacts[ci];
            if (c->IsEnabled() == false) continue;
            if (c->toiCount > kMaxSubSteps) continue;
            float alpha = 1.0f;
            if (c->toiFlag) {
                alpha = c->toi;
            } else {
                Fixture* fA = c->fixtureA;
                Fixture* fB = c->fixtureB;
                if (fA->isSensor || fB->isSensor) continue;
                Body* bA = fA->body;
                Body* bB = fB->body;
                BodyType typeA = bA->type;
                BodyType typeB = bB->type;
                bool activeA = bA->IsAwake() && typeA != kStaticBody;
                bool activeB = bB->IsAwake() && typeB != kStaticBody;
                if (activeA == false && activeB == false) continue;
                bool collideA = typeA != kDynamicBody;  // no bullets in the reference
                bool collideB = typeB != kDynamicBody;
                if (collideA == false && collideB == false) continue;

                float alpha0 = bA->sweep.alpha0;
                if (bA->sweep.alpha0 < bB->sweep.alpha0) {
                    alpha0 = bB->sweep.alpha0;
                    bA->sweep.Advance(alpha0);
                } else if (bB->sweep.alpha0 < bA->sweep.alpha0) {
                    alpha0 = bA->sweep.alpha0;
                    bB->sweep.Advance(alpha0);
                }
                TOIInput input;
                input.proxyA = &fA->shape;
                input.proxyB = &fB->shape;
                input.sweepA = bA->sweep;
                input.sweepB = bB->sweep;
                input.tMax = 1.0f;
                TOIOutput output;
                TimeOfImpact(&output, &input);
                ++stats.toiCalls;
                float beta = output.t;
                if (output.state == TOIOutput::e_touching) alpha = Min(alpha0 + (1.0f - alpha0) * beta, 1.0f);
                else alpha = 1.0f;
                c->toi = alpha;
                c->toiFlag = true;
            }
            if (alpha < minAlpha) {
                minContact = c;
                minAlpha = alpha;
            }
        }

        if (minContact == nullptr || 1.0f - 10.0f * kEpsilon < minAlpha) {
            stepComplete = true;
            break;
        }
        ++stats.toiEvents;

        Fixture* fA = minContact->fixtureA;
        Fixture* fB = minContact->fixtureB;
        Body* bA = fA->body;
        Body* bB = fB->body;
        Sweep backup1 = bA->sweep;
        Sweep backup2 = bB->sweep;
        bA->Advance(minAlpha);
        bB->Advance(minAlpha);

        UpdateContact(minContact);
        minContact->toiFlag = false;
        ++minContact->toiCount;

        if (minContact->IsEnabled() == false || minContact->IsTouching() == false) {
            minContact->enabled = false;
            bA->sweep = backup1;
            bB->sweep = backup2;
            bA->SynchronizeTransform();
            bB->SynchronizeTransform();
            continue;
        }
        bA->SetAwake(true);
        bB->SetAwake(true);

        std::vector<Body*> ibodies;
        std::vector<Contact*> icontacts;
        bA->islandIndex = 0;
        ibodies.push_back(bA);
        bB->islandIndex = 1;
        ibodies.push_back(bB);
        icontacts.push_back(minContact);
        bA->islandFlag = true;
        bB->islandFlag = true;
        minContact->islandFlag = true;

        Body* tbodies[2] = {bA, bB};
        for (int i = 0; i < 2; ++i) {
            Body* body = tbodies[i];
            if (body->type == kDynamicBody) {
                for (int e = (int)body->contacts.size() - 1; e >= 0; --e) {
                    if ((int)ibodies.size() == 2 * kMaxTOIContacts) break;
                    if ((int)icontacts.size() == kMaxTOIContacts) break;
                    Contact* contact = body->contacts[e];
                    if (contact->islandFlag) continue;
                    Body* other = contact->fixtureA->body == body ? contact->fixtureB->body : contact->fixtureA->body;
                    if (other->type == kDynamicBody) continue;  // no bullets
                    if (contact->fixtureA->isSensor || contact->fixtureB->isSensor) continue;
                    Sweep backup = other->sweep;
                    if (other->islandFlag == false) other->Advance(minAlpha);
                    UpdateContact(contact);
                    if (contact->IsEnabled() == false) {
                        other->sweep = backup;
                        other->SynchronizeTransform();
                        continue;
                    }
                    if (contact->IsTouching() == false) {
                        other->sweep = backup;
                        other->SynchronizeTransform();
                        continue;
                    }
                    contact->islandFlag = true;
                    icontacts.push_back(contact);
                    if (other->islandFlag) continue;
                    other->islandFlag = true;
                    if (other->type != kStaticBody) other->SetAwake(true);
                    other->islandIndex = (int)ibodies.size();
                    ibodies.push_back(other);
                }
            }
        }

        TimeStep subStep;
        subStep.dt = (1.0f - minAlpha) * step.dt;
        subStep.inv_dt = 1.0f / subStep.dt;
        subStep.dtRatio = 1.0f;
        subStep.positionIterations = 20;
        subStep.velocityIterations = step.velocityIterations;
        subStep.warmStarting = false;

        // b2Island::SolveTOI
        {
            int toiIndexA = bA->islandIndex, toiIndexB = bB->islandIndex;
            std::vector<Position> positions(ibodies.size());
            std::vector<Velocity> velocities(ibodies.size());
            for (size_t i = 0; i < ibodies.size(); ++i) {
                Body* b = ibodies[i];
                positions[i].c = b->sweep.c;
                positions[i].a = b->sweep.a;
                velocities[i].v = b->linearVelocity;
                velocities[i].w = b->angularVelocity;
            }
            ContactSolver contactSolver(subStep, icontacts, positions.data(), velocities.data());
            for (int i = 0; i < subStep.positionIterations; ++i) {
                bool contactsOkay = contactSolver.SolvePositions(toiIndexA, toiIndexB, true);
                if (contactsOkay) break;
            }
            ibodies[toiIndexA]->sweep.c0 = positions[toiIndexA].c;
            ibodies[toiIndexA]->sweep.a0 = positions[toiIndexA].a;
            ibodies[toiIndexB]->sweep.c0 = positions[toiIndexB].c;
            ibodies[toiIndexB]->sweep.a0 = positions[toiIndexB].a;

            contactSolver.InitializeVelocityConstraints();
            for (int i = 0; i < subStep.velocityIterations; ++i) contactSolver.SolveVelocityConstraints();

            float h = subStep.dt;
            for (size_t i = 0; i < ibodies.size(); ++i) {
                Vec2 c = positions[i].c;
                float a = positions[i].a;
                Vec2 v = velocities[i].v;
                float w = velocities[i].w;
                Vec2 translation = h * v;
                if (Dot(translation, translation) > kMaxTranslationSquared) {
                    float ratio = kMaxTranslation / translation.Length();
                    v *= ratio;
                }
                float rotation = h * w;
                if (rotation * rotation > kMaxRotationSquared) {
                    float ratio = kMaxRotation / Abs(rotation);
                    w *= ratio;
                }
                c += h * v;
                a += h * w;
                positions[i].c = c;
                positions[i].a = a;
                velocities[i].v = v;
                velocities[i].w = w;
                Body* body = ibodies[i];
                body->sweep.c = c;
                body->sweep.a = a;
                body->linearVelocity = v;
                body->angularVelocity = w;
                body->SynchronizeTransform();
            }
        }

        for (Body* body : ibodies) {
            body->islandFlag = false;
            if (body->type != kDynamicBody) continue;
            body->SynchronizeFixtures();
            for (Contact* ce : body->contacts) {
                ce->toiFlag = false;
                ce->islandFlag = false;
            }
        }
        FindNewContacts();
    }
}

void World::Step(float dt, int velocityIterations, int positionIterations) {
    if (newFixture) {
        FindNewContacts();
        newFixture = false;
    }
    TimeStep step;
    step.dt = dt;
    step.velocityIterations = velocityIterations;
    step.positionIterations = positionIterations;
    if (dt > 0.0f) step.inv_dt = 1.0f / dt;
    else step.inv_dt = 0.0f;
    step.dtRatio = inv_dt0 * dt;
    step.warmStarting = warmStarting;
    stats.toiEvents = 0;
    stats.toiCalls = 0;

    Collide();
    if (stepComplete && step.dt > 0.0f) Solve(step);
    if (continuousPhysics && step.dt > 0.0f) SolveTOI(step);
    if (step.dt > 0.0f) inv_dt0 = step.inv_dt;
    ClearForces();  // e_clearForces default
}

}  // namespace b2o
