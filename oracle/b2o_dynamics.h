// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// Restates Box2D 2.3.1 Dynamics/{b2World,b2Body,b2Fixture,b2Island,b2ContactManager}.cpp,
// Dynamics/Contacts/{b2Contact,b2ContactSolver}.cpp and Dynamics/Joints/{b2RevoluteJoint,
// b2FrictionJoint}.cpp as used by /root/reference/src/sim_env/Box2DWorld.cpp (creation sites
// :65,:125,:147,:915,:3655-3661,:3668; stepping :3278-3281; contacts :2898-2907).
// Linked lists of upstream are kept as vectors in creation order; "list order" (newest first)
// is reverse iteration.
#pragma once
#include <vector>

#include "b2o_collision.h"

namespace b2o {

struct Body;
struct Fixture;
struct Joint;
struct Contact;
struct World;

enum BodyType { kStaticBody = 0, kKinematicBody = 1, kDynamicBody = 2 };

struct Filter {
    uint16_t categoryBits = 0x0001;
    uint16_t maskBits = 0xFFFF;
    int16_t groupIndex = 0;
};

struct Fixture {
    Body* body = nullptr;
    PolygonShape shape;
    float density = 0.0f;
    float friction = 0.2f;
    float restitution = 0.0f;
    bool isSensor = false;
    Filter filter;
    // broadphase proxy (one child per polygon)
    AABB fatAABB;     // b2DynamicTree node AABB
    int proxyId = -1; // creation index (monotone in upstream leaf ids)
    void* userData = nullptr;
};

struct JointEdge {
    Joint* joint;
    Body* other;
};

struct Body {
    BodyType type = kStaticBody;
    World* world = nullptr;
    int index = -1;  // creation index
    Transform xf;
    Sweep sweep;
    Vec2 linearVelocity;
    float angularVelocity = 0.0f;
    Vec2 force;
    float torque = 0.0f;
    float mass = 0.0f, invMass = 0.0f;
    float I = 0.0f, invI = 0.0f;
    float sleepTime = 0.0f;
    bool awake = true;
    bool islandFlag = false;
    bool autoSleep = true;
    int islandIndex = 0;
    std::vector<Fixture*> fixtures;      // creation order (upstream list = reverse)
    std::vector<JointEdge> jointEdges;   // creation order (upstream list = reverse)
    std::vector<Contact*> contacts;      // creation order (upstream contact-edge list = reverse)
    void* userData = nullptr;

    Fixture* CreateFixture(const PolygonShape& shape, float density, float friction, float restitution, bool sensor);
    void ResetMassData();
    void GetMassData(MassData* md) const;
    void SetMassData(const MassData* md);
    void SetTransform(const Vec2& position, float angle);
    void SynchronizeFixtures();
    void SynchronizeTransform();
    void Advance(float alpha);
    void SetAwake(bool flag);
    bool IsAwake() const { return awake; }
    bool ShouldCollide(const Body* other) const;
    void SetLinearVelocity(const Vec2& v);
    void SetAngularVelocity(float w);
    void ApplyForceToCenter(const Vec2& f, bool wake);
    void ApplyTorque(float t, bool wake);
    Vec2 GetWorldPoint(const Vec2& p) const { return Mul(xf, p); }
    Vec2 GetWorldCenter() const { return sweep.c; }
    Vec2 GetLocalCenter() const { return sweep.localCenter; }
    float GetAngle() const { return sweep.a; }
    float GetMass() const { return mass; }
    float GetInertia() const { return I + mass * Dot(sweep.localCenter, sweep.localCenter); }
};

struct TimeStep {
    float dt, inv_dt, dtRatio;
    int velocityIterations, positionIterations;
    bool warmStarting;
};

enum JointKind { kFrictionJoint = 0, kRevoluteJoint = 1, kPrismaticJoint = 2 };
enum LimitState { kInactiveLimit = 0, kAtLowerLimit = 1, kAtUpperLimit = 2, kEqualLimits = 3 };

struct Joint {
    JointKind kind;
    Body* bodyA = nullptr;
    Body* bodyB = nullptr;
    bool collideConnected = false;
    bool islandFlag = false;
    int index = -1;  // creation index
    Vec2 localAnchorA, localAnchorB;
    // friction
    Vec2 linearImpulse;
    float angularImpulse = 0.0f;
    float maxForce = 0.0f, maxTorque = 0.0f;
    Mat22 linearMass;
    float angularMass = 0.0f;
    // revolute
    Vec3 impulse;
    float motorImpulse = 0.0f;
    bool enableMotor = false;
    float maxMotorTorque = 0.0f;
    float motorSpeed = 0.0f;
    bool enableLimit = false;
    float referenceAngle = 0.0f;
    float lowerAngle = 0.0f, upperAngle = 0.0f;
    Mat33 massMat;
    float motorMass = 0.0f;
    int limitState = kInactiveLimit;
    // prismatic (b2PrismaticJoint 2.3.1; shares impulse / motorImpulse / enableMotor / motorSpeed / enableLimit /
    // referenceAngle / limitState / massMat / motorMass with the revolute joint's fields)
    Vec2 localXAxisA, localYAxisA;
    float lowerTranslation = 0.0f, upperTranslation = 0.0f;
    float maxMotorForce = 0.0f;
    Vec2 axis, perp;
    float s1 = 0.0f, s2 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    // solver temp
    Vec2 rA, rB, localCenterA, localCenterB;
    float invMassA = 0.0f, invMassB = 0.0f, invIA = 0.0f, invIB = 0.0f;

    void InitVelocityConstraints(const TimeStep& step);
    void SolveVelocityConstraints(const TimeStep& step);
    bool SolvePositionConstraints();
    // revolute API used by the wrapper (Box2DWorld.cpp:972, 1010-1012, 1037, 1211-1218)
    float GetJointAngle() const { return bodyB->sweep.a - bodyA->sweep.a - referenceAngle; }
    float GetJointSpeed() const { return bodyB->angularVelocity - bodyA->angularVelocity; }
    Vec2 GetAnchorA() const { return bodyA->GetWorldPoint(localAnchorA); }
    void EnableMotor(bool flag) { bodyA->SetAwake(true); bodyB->SetAwake(true); enableMotor = flag; }
    void SetMaxMotorTorque(float t) { bodyA->SetAwake(true); bodyB->SetAwake(true); maxMotorTorque = t; }
    void SetMotorSpeed(float s) { bodyA->SetAwake(true); bodyB->SetAwake(true); motorSpeed = s; }
    // prismatic API used by the wrapper (Box2DWorld.cpp:974-977, 1039-1042)
    float GetJointTranslation() const;
    float GetPrismaticJointSpeed() const;
    void PrismaticInit(const TimeStep& step);
    void PrismaticSolveVelocity(const TimeStep& step);
    bool PrismaticSolvePosition();
};

struct Contact {
    Fixture* fixtureA = nullptr;
    Fixture* fixtureB = nullptr;
    Manifold manifold;
    bool touching = false;
    bool enabled = true;
    bool filterFlag = false;
    bool islandFlag = false;
    bool toiFlag = false;
    int toiCount = 0;
    float toi = 1.0f;
    float friction = 0.0f;
    float restitution = 0.0f;
    uint64_t serial = 0;  // creation serial, for diagnostics only
    bool IsTouching() const { return touching; }
    bool IsEnabled() const { return enabled; }
    void GetWorldManifold(WorldManifold* wm) const;
};

struct ContactListener {
    virtual ~ContactListener() {}
    virtual void BeginContact(Contact*) {}
    virtual void EndContact(Contact*) {}
};

struct StepStats {
    int islands = 0;
    int toiEvents = 0;
    int toiCalls = 0;
};

struct World {
    std::vector<Body*> bodies;     // creation order (upstream list = reverse)
    std::vector<Joint*> joints;    // creation order
    std::vector<Contact*> contacts;  // creation order (upstream list = reverse)
    std::vector<Fixture*> proxies;   // proxy id -> fixture
    std::vector<int> moveBuffer;
    Vec2 gravity;
    bool allowSleep = true;
    bool newFixture = false;
    bool warmStarting = true;
    bool continuousPhysics = true;
    bool stepComplete = true;
    float inv_dt0 = 0.0f;
    uint64_t contactSerial = 0;
    ContactListener* listener = nullptr;
    StepStats stats;
    // island trace of the last Step (for ordering parity tests): per island body/contact/joint ids
    bool traceIslands = false;
    std::vector<std::vector<int>> traceBodies, traceJoints;
    std::vector<std::vector<std::pair<int, int>>> traceContacts;

    explicit World(const Vec2& g) : gravity(g) {}
    ~World();
    Body* CreateBody(BodyType type, const Vec2& position, float angle);
    Joint* CreateFrictionJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, float maxForce, float maxTorque);
    Joint* CreateRevoluteJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, float refAngle,
                               float lower, float upper, bool enableLimit, bool enableMotor);
    Joint* CreatePrismaticJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB, const Vec2& axisA, float refAngle,
                                float lower, float upper, bool enableLimit, float maxMotorForce, bool enableMotor);
    void SetAllowSleeping(bool flag);
    void Step(float dt, int velocityIterations, int positionIterations);
    void ClearForces();
    void UpdateContacts();  // fork-only API called at Box2DWorld.cpp:2898; assumed FindNewContacts + Collide

    // contact manager
    void FindNewContacts();
    void Collide();
    void AddPair(Fixture* fA, Fixture* fB);
    void DestroyContact(Contact* c);
    void UpdateContact(Contact* c);
    void MoveProxy(Fixture* f, const AABB& aabb, const Vec2& displacement);
    void Solve(const TimeStep& step);
    void SolveTOI(const TimeStep& step);
};

}  // namespace b2o
