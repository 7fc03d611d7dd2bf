// Region query and feasibility check around the hot path:
//   Box2DWorld::getObjects(aabb, ...)   src/sim_env/Box2DWorld.cpp:3207-3254 with Box2DAABBQuery :2979-2998
//   Box2DWorld::isPhysicallyFeasible    :3370-3400
#pragma once
#include "b2g_toi.cuh"

namespace b2g {

// b2TestOverlap(shapeA, shapeB, xfA, xfB) (b2Collision.cpp): GJK distance with radii < 10 * epsilon
DEV bool testOverlapFixtures(Ctx c, int fA, const XF& xfA, int fB, const XF& xfB) {
    SimplexCache cache;
    cache.count = 0;
    cache.metric = 0.0f;
    float distance = distanceGJK(c, cache, fA, xfA, fB, xfB);
    const float rA = B2G_POLYGON_RADIUS, rB = B2G_POLYGON_RADIUS;
    if (distance > rA + rB && distance > FLT_EPSILON) distance -= rA + rB;
    else distance = 0.0f;
    return distance < 10.0f * FLT_EPSILON;
}

// getObjects(aabb): `fQuery` is the scratch fixture record the kernel prologue filled with the query box
// (b2PolygonShape::SetAsBox(extents)); hits[o] = 1 for every object with a fixture that overlaps it
DEVN void objectsInAABB(Ctx c, float lx, float ly, float ux, float uy, int fQuery, unsigned char* hits) {
    const ModelHeader& h = H();
    for (int o = 0; o < h.no; ++o) hits[o] = 0;
    Box q;
    q.lx = lx; q.ly = ly; q.ux = ux; q.uy = uy;
    XF xfQ;
    xfQ.p = mk(0.5f * (lx + ux), 0.5f * (ly + uy));  // b2AABB::GetCenter
    xfQ.q.s = 0.0f;                                   // b2Rot(0)
    xfQ.q.c = 1.0f;
    for (int f = 0; f < h.nf; ++f) {
        const int b = mI(c, mfix(c, f, MF_BODY));
        if (b == 0) continue;                         // ground: b_collision_ignore (:2988)
        if (!boxOverlap(q, ldFat(c, f))) continue;    // b2DynamicTree::Query on the fat AABBs
        if (!testOverlapFixtures(c, fQuery, xfQ, f, ldXF(c, b))) continue;
        hits[mI(c, mbody(c, b, MB_OBJECT))] = 1;
    }
}

// area of the intersection of two convex polygons given by their world vertices (counter-clockwise):
// Sutherland-Hodgman clipping of A against every edge of B, then the shoelace formula
DEV float convexIntersectionArea(const V2* a, int na, const V2* b, int nb) {
    V2 bufA[2 * 8 + 2], bufB[2 * 8 + 2];
    V2* in = bufA;
    V2* out = bufB;
    int n = na;
    for (int i = 0; i < na; ++i) in[i] = a[i];
    for (int e = 0; e < nb && n > 0; ++e) {
        const V2 c1 = b[e], c2 = b[e + 1 < nb ? e + 1 : 0];
        const V2 edge = c2 - c1;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const V2 p = in[i], q = in[i + 1 < n ? i + 1 : 0];
            const float dp = cross(edge, p - c1), dq = cross(edge, q - c1);
            if (dq >= 0.0f) {
                if (dp < 0.0f) { float t = dp / (dp - dq); out[m++] = p + t * (q - p); }
                out[m++] = q;
            } else if (dp >= 0.0f) {
                float t = dp / (dp - dq);
                out[m++] = p + t * (q - p);
            }
        }
        V2* tmp = in; in = out; out = tmp;
        n = m;
    }
    if (n < 3) return 0.0f;
    float twice = 0.0f;
    for (int i = 1; i + 1 < n; ++i) twice += cross(in[i] - in[0], in[i + 1] - in[0]);
    return 0.5f * twice;
}

// isPhysicallyFeasible: after UpdateContacts(), no pair of polygons of two touching links may overlap by more than
// MAX_ALLOWED_INTERSECTION_AREA = 1e-4 (user units, :22)
DEVN bool physicallyFeasible(Ctx c) {
    findNewContacts(c);
    collide(c);
    const float inv = H().invScale;
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int flags = ldI(c, cslot(c, k, SC_FLAGS));
        if ((flags & CF_TOUCHING) == 0 || (flags & CF_ENABLED) == 0) continue;
        int bA, bB, fA, fB;
        contactBodies(c, k, bA, bB, fA, fB);
        const XF xfA = ldXF(c, bA), xfB = ldXF(c, bB);
        const int sA = mI(c, mbody(c, bA, MB_FIX_START)), nA = mI(c, mbody(c, bA, MB_FIX_COUNT));
        const int sB = mI(c, mbody(c, bB, MB_FIX_START)), nB = mI(c, mbody(c, bB, MB_FIX_COUNT));
        for (int ia = sA; ia < sA + nA; ++ia) {
            V2 va[8];
            const int ca = mI(c, mfix(c, ia, MF_COUNT));
            for (int i = 0; i < ca; ++i) va[i] = mul(xfA, fixVert(c, ia, i));
            for (int ib = sB; ib < sB + nB; ++ib) {
                V2 vb[8];
                const int cb = mI(c, mfix(c, ib, MF_COUNT));
                for (int i = 0; i < cb; ++i) vb[i] = mul(xfB, fixVert(c, ib, i));
                float area = convexIntersectionArea(va, ca, vb, cb);
                if (inv * inv * area > 0.0001f) return false;
            }
        }
    }
    return true;
}

}  // namespace b2g
