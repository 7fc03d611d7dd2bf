// ORACLE — TEST INFRASTRUCTURE ONLY (see b2o_math.h header).  PARITY UNPINNED.
// Box2D 2.3.1 collision routines restated from the published algorithm (SURVEY.md Appendix A.1,
// A.3, A.8, A.9).  Floating-point expression order follows upstream so that a build against the
// real library would agree to the last bit wherever libm is not involved.
#include "b2o_collision.h"

namespace b2o {

// ---- b2PolygonShape ---------------------------------------------------------------------------
static Vec2 ComputeCentroid(const Vec2* vs, int count) {
    Vec2 c(0.0f, 0.0f);
    float area = 0.0f;
    Vec2 pRef(0.0f, 0.0f);
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
        Vec2 p1 = pRef;
        Vec2 p2 = vs[i];
        Vec2 p3 = i + 1 < count ? vs[i + 1] : vs[0];
        Vec2 e1 = p2 - p1;
        Vec2 e2 = p3 - p1;
        float D = Cross(e1, e2);
        float triangleArea = 0.5f * D;
        area += triangleArea;
        c += triangleArea * inv3 * (p1 + p2 + p3);
    }
    c *= 1.0f / area;
    return c;
}

void PolygonShape::SetAsBox(float hx, float hy) {
    count = 4;
    vertices[0].Set(-hx, -hy);
    vertices[1].Set(hx, -hy);
    vertices[2].Set(hx, hy);
    vertices[3].Set(-hx, hy);
    normals[0].Set(0.0f, -1.0f);
    normals[1].Set(1.0f, 0.0f);
    normals[2].Set(0.0f, 1.0f);
    normals[3].Set(-1.0f, 0.0f);
    centroid.SetZero();
}

void PolygonShape::Set(const Vec2* points, int cnt) {
    if (cnt < 3) { SetAsBox(1.0f, 1.0f); return; }
    int n = Min(cnt, kMaxPolygonVertices);
    // weld
    Vec2 ps[kMaxPolygonVertices];
    int tempCount = 0;
    for (int i = 0; i < n; ++i) {
        Vec2 v = points[i];
        bool unique = true;
        for (int j = 0; j < tempCount; ++j) {
            if (DistanceSquared(v, ps[j]) < 0.5f * kLinearSlop) { unique = false; break; }
        }
        if (unique) ps[tempCount++] = v;
    }
    n = tempCount;
    if (n < 3) { SetAsBox(1.0f, 1.0f); return; }
    // right-most point, ties to lower y
    int i0 = 0;
    float x0 = ps[0].x;
    for (int i = 1; i < n; ++i) {
        float x = ps[i].x;
        if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
    }
    int hull[kMaxPolygonVertices];
    int m = 0;
    int ih = i0;
    for (;;) {
        hull[m] = ih;
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            Vec2 r = ps[ie] - ps[hull[m]];
            Vec2 v = ps[j] - ps[hull[m]];
            float c = Cross(r, v);
            if (c < 0.0f) ie = j;
            if (c == 0.0f && v.LengthSquared() > r.LengthSquared()) ie = j;
        }
        ++m;
        ih = ie;
        if (ie == i0) break;
    }
    if (m < 3) { SetAsBox(1.0f, 1.0f); return; }
    count = m;
    for (int i = 0; i < m; ++i) vertices[i] = ps[hull[i]];
    for (int i = 0; i < m; ++i) {
        int i1 = i;
        int i2 = i + 1 < m ? i + 1 : 0;
        Vec2 edge = vertices[i2] - vertices[i1];
        normals[i] = Cross(edge, 1.0f);
        normals[i].Normalize();
    }
    centroid = ComputeCentroid(vertices, m);
}

void PolygonShape::ComputeMass(MassData* massData, float density) const {
    Vec2 center(0.0f, 0.0f);
    float area = 0.0f;
    float I = 0.0f;
    Vec2 s(0.0f, 0.0f);
    for (int i = 0; i < count; ++i) s += vertices[i];
    s *= 1.0f / count;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
        Vec2 e1 = vertices[i] - s;
        Vec2 e2 = i + 1 < count ? vertices[i + 1] - s : vertices[0] - s;
        float D = Cross(e1, e2);
        float triangleArea = 0.5f * D;
        area += triangleArea;
        center += triangleArea * k_inv3 * (e1 + e2);
        float ex1 = e1.x, ey1 = e1.y;
        float ex2 = e2.x, ey2 = e2.y;
        float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
        float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    massData->mass = density * area;
    center *= 1.0f / area;
    massData->center = center + s;
    massData->I = density * I;
    massData->I += massData->mass * (Dot(massData->center, massData->center) - Dot(center, center));
}

void PolygonShape::ComputeAABB(AABB* aabb, const Transform& xf) const {
    Vec2 lower = Mul(xf, vertices[0]);
    Vec2 upper = lower;
    for (int i = 1; i < count; ++i) {
        Vec2 v = Mul(xf, vertices[i]);
        lower = Min(lower, v);
        upper = Max(upper, v);
    }
    Vec2 r(radius, radius);
    aabb->lowerBound = lower - r;
    aabb->upperBound = upper + r;
}

int PolygonShape::GetSupport(const Vec2& d) const {
    int bestIndex = 0;
    float bestValue = Dot(vertices[0], d);
    for (int i = 1; i < count; ++i) {
        float value = Dot(vertices[i], d);
        if (value > bestValue) { bestIndex = i; bestValue = value; }
    }
    return bestIndex;
}

// ---- b2WorldManifold ----------------------------------------------------------------------------
void WorldManifold::Initialize(const Manifold* manifold, const Transform& xfA, float radiusA,
                               const Transform& xfB, float radiusB) {
    if (manifold->pointCount == 0) return;
    switch (manifold->type) {
    case Manifold::e_faceA: {
        normal = Mul(xfA.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfA, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
            Vec2 clipPoint = Mul(xfB, manifold->points[i].localPoint);
            Vec2 cA = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 cB = clipPoint - radiusB * normal;
            points[i] = 0.5f * (cA + cB);
            separations[i] = Dot(cB - cA, normal);
        }
    } break;
    case Manifold::e_faceB: {
        normal = Mul(xfB.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfB, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
            Vec2 clipPoint = Mul(xfA, manifold->points[i].localPoint);
            Vec2 cB = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 cA = clipPoint - radiusA * normal;
            points[i] = 0.5f * (cA + cB);
            separations[i] = Dot(cA - cB, normal);
        }
        normal = -normal;
    } break;
    default: break;  // circles never occur: the reference only creates polygons
    }
}

// ---- b2CollidePolygons (2.3.1) --------------------------------------------------------------------
struct ClipVertex {
    Vec2 v;
    ContactID id;
};

static int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], const Vec2& normal,
                             float offset, int vertexIndexA) {
    int numOut = 0;
    float distance0 = Dot(normal, vIn[0].v) - offset;
    float distance1 = Dot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        vOut[numOut].id.cf.indexA = (uint8_t)vertexIndexA;
        vOut[numOut].id.cf.indexB = vIn[0].id.cf.indexB;
        vOut[numOut].id.cf.typeA = kFeatureVertex;
        vOut[numOut].id.cf.typeB = kFeatureFace;
        ++numOut;
    }
    return numOut;
}

static float FindMaxSeparation(int* edgeIndex, const PolygonShape* poly1, const Transform& xf1,
                               const PolygonShape* poly2, const Transform& xf2) {
    int count1 = poly1->count;
    int count2 = poly2->count;
    const Vec2* n1s = poly1->normals;
    const Vec2* v1s = poly1->vertices;
    const Vec2* v2s = poly2->vertices;
    Transform xf = MulT(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -kMaxFloat;
    for (int i = 0; i < count1; ++i) {
        Vec2 n = Mul(xf.q, n1s[i]);
        Vec2 v1 = Mul(xf, v1s[i]);
        float si = kMaxFloat;
        for (int j = 0; j < count2; ++j) {
            float sij = Dot(n, v2s[j] - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

static void FindIncidentEdge(ClipVertex c[2], const PolygonShape* poly1, const Transform& xf1, int edge1,
                             const PolygonShape* poly2, const Transform& xf2) {
    const Vec2* normals1 = poly1->normals;
    int count2 = poly2->count;
    const Vec2* vertices2 = poly2->vertices;
    const Vec2* normals2 = poly2->normals;
    Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, normals1[edge1]));
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < count2; ++i) {
        float dot = Dot(normal1, normals2[i]);
        if (dot < minDot) { minDot = dot; index = i; }
    }
    int i1 = index;
    int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
    c[0].v = Mul(xf2, vertices2[i1]);
    c[0].id.cf.indexA = (uint8_t)edge1;
    c[0].id.cf.indexB = (uint8_t)i1;
    c[0].id.cf.typeA = kFeatureFace;
    c[0].id.cf.typeB = kFeatureVertex;
    c[1].v = Mul(xf2, vertices2[i2]);
    c[1].id.cf.indexA = (uint8_t)edge1;
    c[1].id.cf.indexB = (uint8_t)i2;
    c[1].id.cf.typeA = kFeatureFace;
    c[1].id.cf.typeB = kFeatureVertex;
}

void CollidePolygons(Manifold* manifold, const PolygonShape* polyA, const Transform& xfA,
                     const PolygonShape* polyB, const Transform& xfB) {
    manifold->pointCount = 0;
    float totalRadius = polyA->radius + polyB->radius;

    int edgeA = 0;
    float separationA = FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;

    int edgeB = 0;
    float separationB = FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;

    const PolygonShape* poly1;
    const PolygonShape* poly2;
    Transform xf1, xf2;
    int edge1;
    uint8_t flip;
    const float k_tol = 0.1f * kLinearSlop;

    if (separationB > separationA + k_tol) {
        poly1 = polyB; poly2 = polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB;
        manifold->type = Manifold::e_faceB;
        flip = 1;
    } else {
        poly1 = polyA; poly2 = polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA;
        manifold->type = Manifold::e_faceA;
        flip = 0;
    }

    ClipVertex incidentEdge[2];
    FindIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);

    int count1 = poly1->count;
    const Vec2* vertices1 = poly1->vertices;
    int iv1 = edge1;
    int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    Vec2 v11 = vertices1[iv1];
    Vec2 v12 = vertices1[iv2];

    Vec2 localTangent = v12 - v11;
    localTangent.Normalize();
    Vec2 localNormal = Cross(localTangent, 1.0f);
    Vec2 planePoint = 0.5f * (v11 + v12);

    Vec2 tangent = Mul(xf1.q, localTangent);
    Vec2 normal = Cross(tangent, 1.0f);

    v11 = Mul(xf1, v11);
    v12 = Mul(xf1, v12);

    float frontOffset = Dot(normal, v11);
    float sideOffset1 = -Dot(tangent, v11) + totalRadius;
    float sideOffset2 = Dot(tangent, v12) + totalRadius;

    ClipVertex clipPoints1[2];
    ClipVertex clipPoints2[2];
    int np;
    np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;

    manifold->localNormal = localNormal;
    manifold->localPoint = planePoint;

    int pointCount = 0;
    for (int i = 0; i < kMaxManifoldPoints; ++i) {
        float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            ManifoldPoint* cp = manifold->points + pointCount;
            cp->localPoint = MulT(xf2, clipPoints2[i].v);
            cp->id = clipPoints2[i].id;
            if (flip) {
                ContactID cf = cp->id;
                cp->id.cf.indexA = cf.cf.indexB;
                cp->id.cf.indexB = cf.cf.indexA;
                cp->id.cf.typeA = cf.cf.typeB;
                cp->id.cf.typeB = cf.cf.typeA;
            }
            ++pointCount;
        }
    }
    manifold->pointCount = pointCount;
}

// ---- b2Distance (GJK) -----------------------------------------------------------------------------
namespace {
struct SimplexVertex {
    Vec2 wA, wB, w;
    float a;
    int indexA, indexB;
};

struct Simplex {
    SimplexVertex v[3];
    int count;

    void ReadCache(const SimplexCache* cache, const PolygonShape* proxyA, const Transform& transformA,
                   const PolygonShape* proxyB, const Transform& transformB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* sv = v + i;
            sv->indexA = cache->indexA[i];
            sv->indexB = cache->indexB[i];
            Vec2 wALocal = proxyA->vertices[sv->indexA];
            Vec2 wBLocal = proxyB->vertices[sv->indexB];
            sv->wA = Mul(transformA, wALocal);
            sv->wB = Mul(transformB, wBLocal);
            sv->w = sv->wB - sv->wA;
            sv->a = 0.0f;
        }
        if (count > 1) {
            float metric1 = cache->metric;
            float metric2 = GetMetric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
        }
        if (count == 0) {
            SimplexVertex* sv = v + 0;
            sv->indexA = 0;
            sv->indexB = 0;
            Vec2 wALocal = proxyA->vertices[0];
            Vec2 wBLocal = proxyB->vertices[0];
            sv->wA = Mul(transformA, wALocal);
            sv->wB = Mul(transformB, wBLocal);
            sv->w = sv->wB - sv->wA;
            sv->a = 1.0f;
            count = 1;
        }
    }
    void WriteCache(SimplexCache* cache) const {
        cache->metric = GetMetric();
        cache->count = (uint16_t)count;
        for (int i = 0; i < count; ++i) {
            cache->indexA[i] = (uint8_t)v[i].indexA;
            cache->indexB[i] = (uint8_t)v[i].indexB;
        }
    }
    Vec2 GetSearchDirection() const {
        switch (count) {
        case 1: return -v[0].w;
        case 2: {
            Vec2 e12 = v[1].w - v[0].w;
            float sgn = Cross(e12, -v[0].w);
            if (sgn > 0.0f) return Cross(1.0f, e12);
            else return Cross(e12, 1.0f);
        }
        default: return Vec2(0.0f, 0.0f);
        }
    }
    Vec2 GetClosestPoint() const {
        switch (count) {
        case 1: return v[0].w;
        case 2: return v[0].a * v[0].w + v[1].a * v[1].w;
        default: return Vec2(0.0f, 0.0f);
        }
    }
    void GetWitnessPoints(Vec2* pA, Vec2* pB) const {
        switch (count) {
        case 1: *pA = v[0].wA; *pB = v[0].wB; break;
        case 2:
            *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
            *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
            break;
        case 3:
            *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
            *pB = *pA;
            break;
        default: break;
        }
    }
    float GetMetric() const {
        switch (count) {
        case 1: return 0.0f;
        case 2: return Distance(v[0].w, v[1].w);
        case 3: return Cross(v[1].w - v[0].w, v[2].w - v[0].w);
        default: return 0.0f;
        }
    }
    void Solve2() {
        Vec2 w1 = v[0].w;
        Vec2 w2 = v[1].w;
        Vec2 e12 = w2 - w1;
        float d12_2 = -Dot(w1, e12);
        if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        float d12_1 = Dot(w2, e12);
        if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        float inv_d12 = 1.0f / (d12_1 + d12_2);
        v[0].a = d12_1 * inv_d12;
        v[1].a = d12_2 * inv_d12;
        count = 2;
    }
    void Solve3() {
        Vec2 w1 = v[0].w;
        Vec2 w2 = v[1].w;
        Vec2 w3 = v[2].w;
        Vec2 e12 = w2 - w1;
        float w1e12 = Dot(w1, e12);
        float w2e12 = Dot(w2, e12);
        float d12_1 = w2e12;
        float d12_2 = -w1e12;
        Vec2 e13 = w3 - w1;
        float w1e13 = Dot(w1, e13);
        float w3e13 = Dot(w3, e13);
        float d13_1 = w3e13;
        float d13_2 = -w1e13;
        Vec2 e23 = w3 - w2;
        float w2e23 = Dot(w2, e23);
        float w3e23 = Dot(w3, e23);
        float d23_1 = w3e23;
        float d23_2 = -w2e23;
        float n123 = Cross(e12, e13);
        float d123_1 = n123 * Cross(w2, w3);
        float d123_2 = n123 * Cross(w3, w1);
        float d123_3 = n123 * Cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            float inv_d12 = 1.0f / (d12_1 + d12_2);
            v[0].a = d12_1 * inv_d12;
            v[1].a = d12_2 * inv_d12;
            count = 2;
            return;
        }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            float inv_d13 = 1.0f / (d13_1 + d13_2);
            v[0].a = d13_1 * inv_d13;
            v[2].a = d13_2 * inv_d13;
            count = 2;
            v[1] = v[2];
            return;
        }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            float inv_d23 = 1.0f / (d23_1 + d23_2);
            v[1].a = d23_1 * inv_d23;
            v[2].a = d23_2 * inv_d23;
            count = 2;
            v[0] = v[2];
            return;
        }
        float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv_d123;
        v[1].a = d123_2 * inv_d123;
        v[2].a = d123_3 * inv_d123;
        count = 3;
    }
};
}  // namespace

void DistanceGJK(DistanceOutput* output, SimplexCache* cache, const DistanceInput* input) {
    const PolygonShape* proxyA = input->proxyA;
    const PolygonShape* proxyB = input->proxyB;
    Transform transformA = input->transformA;
    Transform transformB = input->transformB;

    Simplex simplex;
    simplex.ReadCache(cache, proxyA, transformA, proxyB, transformB);
    SimplexVertex* vertices = simplex.v;
    const int k_maxIters = 20;
    int saveA[3], saveB[3];
    int saveCount = 0;

    int iter = 0;
    while (iter < k_maxIters) {
        saveCount = simplex.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = vertices[i].indexA; saveB[i] = vertices[i].indexB; }
        switch (simplex.count) {
        case 1: break;
        case 2: simplex.Solve2(); break;
        case 3: simplex.Solve3(); break;
        }
        if (simplex.count == 3) break;
        Vec2 d = simplex.GetSearchDirection();
        if (d.LengthSquared() < kEpsilon * kEpsilon) break;
        SimplexVertex* vertex = vertices + simplex.count;
        vertex->indexA = proxyA->GetSupport(MulT(transformA.q, -d));
        vertex->wA = Mul(transformA, proxyA->vertices[vertex->indexA]);
        vertex->indexB = proxyB->GetSupport(MulT(transformB.q, d));
        vertex->wB = Mul(transformB, proxyB->vertices[vertex->indexB]);
        vertex->w = vertex->wB - vertex->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i) {
            if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) { duplicate = true; break; }
        }
        if (duplicate) break;
        ++simplex.count;
    }

    simplex.GetWitnessPoints(&output->pointA, &output->pointB);
    output->distance = Distance(output->pointA, output->pointB);
    output->iterations = iter;
    simplex.WriteCache(cache);

    if (input->useRadii) {
        float rA = proxyA->radius;
        float rB = proxyB->radius;
        if (output->distance > rA + rB && output->distance > kEpsilon) {
            output->distance -= rA + rB;
            Vec2 normal = output->pointB - output->pointA;
            normal.Normalize();
            output->pointA += rA * normal;
            output->pointB -= rB * normal;
        } else {
            Vec2 p = 0.5f * (output->pointA + output->pointB);
            output->pointA = p;
            output->pointB = p;
            output->distance = 0.0f;
        }
    }
}

bool TestOverlapShapes(const PolygonShape* a, const PolygonShape* b, const Transform& xfA, const Transform& xfB) {
    DistanceInput input;
    input.proxyA = a;
    input.proxyB = b;
    input.transformA = xfA;
    input.transformB = xfB;
    input.useRadii = true;
    SimplexCache cache;
    cache.count = 0;
    DistanceOutput output;
    DistanceGJK(&output, &cache, &input);
    return output.distance < 10.0f * kEpsilon;
}

// ---- b2TimeOfImpact ---------------------------------------------------------------------------------
namespace {
struct SeparationFunction {
    enum Type { e_points, e_faceA, e_faceB };
    const PolygonShape* proxyA;
    const PolygonShape* proxyB;
    Sweep sweepA, sweepB;
    Type type;
    Vec2 localPoint;
    Vec2 axis;

    float Initialize(const SimplexCache* cache, const PolygonShape* pA, const Sweep& sA,
                     const PolygonShape* pB, const Sweep& sB, float t1) {
        proxyA = pA;
        proxyB = pB;
        int count = cache->count;
        sweepA = sA;
        sweepB = sB;
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t1);
        sweepB.GetTransform(&xfB, t1);
        if (count == 1) {
            type = e_points;
            Vec2 localPointA = proxyA->vertices[cache->indexA[0]];
            Vec2 localPointB = proxyB->vertices[cache->indexB[0]];
            Vec2 pointA = Mul(xfA, localPointA);
            Vec2 pointB = Mul(xfB, localPointB);
            axis = pointB - pointA;
            float s = axis.Normalize();
            return s;
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = e_faceB;
            Vec2 localPointB1 = proxyB->vertices[cache->indexB[0]];
            Vec2 localPointB2 = proxyB->vertices[cache->indexB[1]];
            axis = Cross(localPointB2 - localPointB1, 1.0f);
            axis.Normalize();
            Vec2 normal = Mul(xfB.q, axis);
            localPoint = 0.5f * (localPointB1 + localPointB2);
            Vec2 pointB = Mul(xfB, localPoint);
            Vec2 localPointA = proxyA->vertices[cache->indexA[0]];
            Vec2 pointA = Mul(xfA, localPointA);
            float s = Dot(pointA - pointB, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        } else {
            type = e_faceA;
            Vec2 localPointA1 = proxyA->vertices[cache->indexA[0]];
            Vec2 localPointA2 = proxyA->vertices[cache->indexA[1]];
            axis = Cross(localPointA2 - localPointA1, 1.0f);
            axis.Normalize();
            Vec2 normal = Mul(xfA.q, axis);
            localPoint = 0.5f * (localPointA1 + localPointA2);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 localPointB = proxyB->vertices[cache->indexB[0]];
            Vec2 pointB = Mul(xfB, localPointB);
            float s = Dot(pointB - pointA, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        }
    }

    float FindMinSeparation(int* indexA, int* indexB, float t) const {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t);
        sweepB.GetTransform(&xfB, t);
        switch (type) {
        case e_points: {
            Vec2 axisA = MulT(xfA.q, axis);
            Vec2 axisB = MulT(xfB.q, -axis);
            *indexA = proxyA->GetSupport(axisA);
            *indexB = proxyB->GetSupport(axisB);
            Vec2 pointA = Mul(xfA, proxyA->vertices[*indexA]);
            Vec2 pointB = Mul(xfB, proxyB->vertices[*indexB]);
            return Dot(pointB - pointA, axis);
        }
        case e_faceA: {
            Vec2 normal = Mul(xfA.q, axis);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 axisB = MulT(xfB.q, -normal);
            *indexA = -1;
            *indexB = proxyB->GetSupport(axisB);
            Vec2 pointB = Mul(xfB, proxyB->vertices[*indexB]);
            return Dot(pointB - pointA, normal);
        }
        case e_faceB: {
            Vec2 normal = Mul(xfB.q, axis);
            Vec2 pointB = Mul(xfB, localPoint);
            Vec2 axisA = MulT(xfA.q, -normal);
            *indexB = -1;
            *indexA = proxyA->GetSupport(axisA);
            Vec2 pointA = Mul(xfA, proxyA->vertices[*indexA]);
            return Dot(pointA - pointB, normal);
        }
        }
        *indexA = -1;
        *indexB = -1;
        return 0.0f;
    }

    float Evaluate(int indexA, int indexB, float t) const {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t);
        sweepB.GetTransform(&xfB, t);
        switch (type) {
        case e_points: {
            Vec2 pointA = Mul(xfA, proxyA->vertices[indexA]);
            Vec2 pointB = Mul(xfB, proxyB->vertices[indexB]);
            return Dot(pointB - pointA, axis);
        }
        case e_faceA: {
            Vec2 normal = Mul(xfA.q, axis);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 pointB = Mul(xfB, proxyB->vertices[indexB]);
            return Dot(pointB - pointA, normal);
        }
        case e_faceB: {
            Vec2 normal = Mul(xfB.q, axis);
            Vec2 pointB = Mul(xfB, localPoint);
            Vec2 pointA = Mul(xfA, proxyA->vertices[indexA]);
            return Dot(pointA - pointB, normal);
        }
        }
        return 0.0f;
    }
};
}  // namespace

void TimeOfImpact(TOIOutput* output, const TOIInput* input) {
    output->state = TOIOutput::e_unknown;
    output->t = input->tMax;

    const PolygonShape* proxyA = input->proxyA;
    const PolygonShape* proxyB = input->proxyB;
    Sweep sweepA = input->sweepA;
    Sweep sweepB = input->sweepB;
    sweepA.Normalize();
    sweepB.Normalize();

    float tMax = input->tMax;
    float totalRadius = proxyA->radius + proxyB->radius;
    float target = Max(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    float tolerance = 0.25f * kLinearSlop;

    float t1 = 0.0f;
    const int k_maxIterations = 20;
    int iter = 0;

    SimplexCache cache;
    cache.count = 0;
    DistanceInput distanceInput;
    distanceInput.proxyA = proxyA;
    distanceInput.proxyB = proxyB;
    distanceInput.useRadii = false;

    for (;;) {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t1);
        sweepB.GetTransform(&xfB, t1);
        distanceInput.transformA = xfA;
        distanceInput.transformB = xfB;
        DistanceOutput distanceOutput;
        DistanceGJK(&distanceOutput, &cache, &distanceInput);

        if (distanceOutput.distance <= 0.0f) {
            output->state = TOIOutput::e_overlapped;
            output->t = 0.0f;
            break;
        }
        if (distanceOutput.distance < target + tolerance) {
            output->state = TOIOutput::e_touching;
            output->t = t1;
            break;
        }

        SeparationFunction fcn;
        fcn.Initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);

        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = fcn.FindMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) {
                output->state = TOIOutput::e_separated;
                output->t = tMax;
                done = true;
                break;
            }
            if (s2 > target - tolerance) {
                t1 = t2;
                break;
            }
            float s1 = fcn.Evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) {
                output->state = TOIOutput::e_failed;
                output->t = t1;
                done = true;
                break;
            }
            if (s1 <= target + tolerance) {
                output->state = TOIOutput::e_touching;
                output->t = t1;
                done = true;
                break;
            }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIterCount;
                float s = fcn.Evaluate(indexA, indexB, t);
                if (Abs(s - target) < tolerance) {
                    t2 = t;
                    break;
                }
                if (s > target) { a1 = t; s1 = s; }
                else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == k_maxIterations) {
            output->state = TOIOutput::e_failed;
            output->t = t1;
            break;
        }
    }
}

}  // namespace b2o
