// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// Restates Box2D 2.3.1 Collision/{Shapes/b2PolygonShape,b2CollidePolygon,b2Collision,b2Distance,
// b2TimeOfImpact}.cpp — the arithmetic behind the reference's only shape type
// (b2PolygonShape, /root/reference/src/sim_env/Box2DWorld.cpp:79,103) and behind b2World::Step
// (Box2DWorld.cpp:3279).
#pragma once
#include "b2o_math.h"

namespace b2o {

struct MassData {
    float mass = 0.0f;
    Vec2 center;
    float I = 0.0f;
};

struct PolygonShape {
    Vec2 centroid;
    Vec2 vertices[kMaxPolygonVertices];
    Vec2 normals[kMaxPolygonVertices];
    int count = 0;
    float radius = kPolygonRadius;

    void Set(const Vec2* points, int n);          // b2PolygonShape::Set (welding + gift wrapping)
    void SetAsBox(float hx, float hy);
    void ComputeMass(MassData* md, float density) const;
    void ComputeAABB(AABB* aabb, const Transform& xf) const;
    int GetSupport(const Vec2& d) const;
};

union ContactID {
    struct { uint8_t indexA, indexB, typeA, typeB; } cf;
    uint32_t key;
};
enum { kFeatureVertex = 0, kFeatureFace = 1 };

struct ManifoldPoint {
    Vec2 localPoint;
    float normalImpulse = 0.0f;
    float tangentImpulse = 0.0f;
    ContactID id;
    ManifoldPoint() { id.key = 0; }
};

struct Manifold {
    enum Type { e_circles = 0, e_faceA = 1, e_faceB = 2 };
    ManifoldPoint points[kMaxManifoldPoints];
    Vec2 localNormal;
    Vec2 localPoint;
    int type = e_circles;
    int pointCount = 0;
};

struct WorldManifold {
    Vec2 normal;
    Vec2 points[kMaxManifoldPoints];
    float separations[kMaxManifoldPoints];
    void Initialize(const Manifold* manifold, const Transform& xfA, float radiusA,
                    const Transform& xfB, float radiusB);
};

void CollidePolygons(Manifold* manifold, const PolygonShape* polyA, const Transform& xfA,
                     const PolygonShape* polyB, const Transform& xfB);

// ---- b2Distance ---------------------------------------------------------------------------------
struct SimplexCache {
    float metric = 0.0f;
    uint16_t count = 0;
    uint8_t indexA[3];
    uint8_t indexB[3];
};
struct DistanceInput {
    const PolygonShape* proxyA;
    const PolygonShape* proxyB;
    Transform transformA, transformB;
    bool useRadii;
};
struct DistanceOutput {
    Vec2 pointA, pointB;
    float distance;
    int iterations;
};
void DistanceGJK(DistanceOutput* output, SimplexCache* cache, const DistanceInput* input);
bool TestOverlapShapes(const PolygonShape* a, const PolygonShape* b, const Transform& xfA, const Transform& xfB);

// ---- b2TimeOfImpact -----------------------------------------------------------------------------
struct TOIInput {
    const PolygonShape* proxyA;
    const PolygonShape* proxyB;
    Sweep sweepA, sweepB;
    float tMax;
};
struct TOIOutput {
    enum State { e_unknown, e_failed, e_overlapped, e_touching, e_separated };
    State state;
    float t;
};
void TimeOfImpact(TOIOutput* output, const TOIInput* input);

}  // namespace b2o
