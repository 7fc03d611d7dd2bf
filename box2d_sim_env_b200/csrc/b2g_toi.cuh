// Continuous collision: b2Distance (GJK), b2TimeOfImpact (conservative advancement) and b2World::SolveTOI with
// its sub-step island (Box2D 2.3.1).  Active in the reference because the wrapper
// never disables continuous physics and static objects exist (test_data/test_env.yaml:10-12): every contact
// between a dynamic and a non-dynamic body is swept each step.
#pragma once
#include "b2g_solver.cuh"

namespace b2g {

struct Sweep {
    V2 lc, c0, c;
    float a0, a, alpha0;
    float keep;  // 4th word of the {c, a} chunk (the body's sleep time): rides along untouched
};

DEV Sweep ldSweep(Ctx c, int b) {
    Sweep s;
    s.lc = localCenter(c, b);
    float4 p = ld4(c, bslot(c, b, SB_POS4)), p0 = ld4(c, bslot(c, b, SB_POS04));
    s.c0 = mk(p0.x, p0.y);
    s.c = mk(p.x, p.y);
    s.a0 = p0.z;
    s.a = p.z;
    s.alpha0 = p0.w;
    s.keep = p.w;
    return s;
}
DEV void stSweep(Ctx c, int b, const Sweep& s) {
    st4(c, bslot(c, b, SB_POS4), mk4(s.c.x, s.c.y, s.a, s.keep));
    st4(c, bslot(c, b, SB_POS04), mk4(s.c0.x, s.c0.y, s.a0, s.alpha0));
}
DEV void sweepAdvance(Sweep& s, float alpha) {  // b2Sweep::Advance
    float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
    s.c0 = s.c0 + beta * (s.c - s.c0);
    s.a0 += beta * (s.a - s.a0);
    s.alpha0 = alpha;
}
DEV XF sweepXF(const Sweep& s, float beta) {  // b2Sweep::GetTransform
    XF xf;
    xf.p = (1.0f - beta) * s.c0 + beta * s.c;
    float angle = (1.0f - beta) * s.a0 + beta * s.a;
    xf.q = rotOf(angle);
    xf.p = xf.p - mul(xf.q, s.lc);
    return xf;
}
DEV void sweepNormalize(Sweep& s) {  // b2Sweep::Normalize
    float twoPi = 2.0f * B2G_PI;
    float d = twoPi * floorf(s.a0 / twoPi);
    s.a0 -= d;
    s.a -= d;
}
// b2Body::Advance
DEV void bodyAdvance(Ctx c, int b, float alpha) {
    Sweep s = ldSweep(c, b);
    sweepAdvance(s, alpha);
    s.c = s.c0;
    s.a = s.a0;
    stSweep(c, b, s);
    syncTransform(c, b);
}

DEV int fixSupport(Ctx c, int f, V2 d) {  // b2DistanceProxy::GetSupport
    int n = mI(c, mfix(c, f, MF_COUNT));
    int best = 0;
    float bestValue = dot(fixVert(c, f, 0), d);
    for (int i = 1; i < n; ++i) {
        float value = dot(fixVert(c, f, i), d);
        if (value > bestValue) { best = i; bestValue = value; }
    }
    return best;
}

// ---- b2Distance -----------------------------------------------------------------------------------------------
struct SimplexCache {
    float metric;
    int count;
    int indexA[3], indexB[3];
};
struct SVertex {
    V2 wA, wB, w;
    float a;
    int indexA, indexB;
};
struct Simplex {
    SVertex v[3];
    int count;
};

DEV float simplexMetric(const Simplex& s) {
    if (s.count == 2) return length(s.v[0].w - s.v[1].w);
    if (s.count == 3) return cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
    return 0.0f;
}

DEV void simplexSolve2(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w;
    V2 e12 = w2 - w1;
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv_d12;
    s.v[1].a = d12_2 * inv_d12;
    s.count = 2;
}

DEV void simplexSolve3(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
    V2 e12 = w2 - w1;
    float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
    float d12_1 = w2e12, d12_2 = -w1e12;
    V2 e13 = w3 - w1;
    float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
    float d13_1 = w3e13, d13_2 = -w1e13;
    V2 e23 = w3 - w2;
    float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
    float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = cross(e12, e13);
    float d123_1 = n123 * cross(w2, w3);
    float d123_2 = n123 * cross(w3, w1);
    float d123_3 = n123 * cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
        float inv = 1.0f / (d12_1 + d12_2);
        s.v[0].a = d12_1 * inv; s.v[1].a = d12_2 * inv; s.count = 2;
        return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
        float inv = 1.0f / (d13_1 + d13_2);
        s.v[0].a = d13_1 * inv; s.v[2].a = d13_2 * inv; s.count = 2; s.v[1] = s.v[2];
        return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s.v[2].a = 1.0f; s.count = 1; s.v[0] = s.v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
        float inv = 1.0f / (d23_1 + d23_2);
        s.v[1].a = d23_1 * inv; s.v[2].a = d23_2 * inv; s.count = 2; s.v[0] = s.v[2];
        return;
    }
    float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    s.v[0].a = d123_1 * inv; s.v[1].a = d123_2 * inv; s.v[2].a = d123_3 * inv; s.count = 3;
}

// b2Distance without radii (the only mode b2TimeOfImpact uses); returns the distance, updates the cache
DEVN float distanceGJK(Ctx c, SimplexCache& cache, int fA, const XF& xfA, int fB, const XF& xfB) {
    Simplex s;
    // ReadCache
    s.count = cache.count;
    for (int i = 0; i < s.count; ++i) {
        SVertex& v = s.v[i];
        v.indexA = cache.indexA[i];
        v.indexB = cache.indexB[i];
        v.wA = mul(xfA, fixVert(c, fA, v.indexA));
        v.wB = mul(xfB, fixVert(c, fB, v.indexB));
        v.w = v.wB - v.wA;
        v.a = 0.0f;
    }
    if (s.count > 1) {
        float metric1 = cache.metric;
        float metric2 = simplexMetric(s);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < FLT_EPSILON) s.count = 0;
    }
    if (s.count == 0) {
        SVertex& v = s.v[0];
        v.indexA = 0;
        v.indexB = 0;
        v.wA = mul(xfA, fixVert(c, fA, 0));
        v.wB = mul(xfB, fixVert(c, fB, 0));
        v.w = v.wB - v.wA;
        v.a = 1.0f;
        s.count = 1;
    }
    int saveA[3], saveB[3];
    int iter = 0;
    while (iter < 20) {
        int saveCount = s.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = s.v[i].indexA; saveB[i] = s.v[i].indexB; }
        if (s.count == 2) simplexSolve2(s);
        else if (s.count == 3) simplexSolve3(s);
        if (s.count == 3) break;
        // GetSearchDirection
        V2 d;
        if (s.count == 1) {
            d = -s.v[0].w;
        } else {
            V2 e12 = s.v[1].w - s.v[0].w;
            float sgn = cross(e12, -s.v[0].w);
            if (sgn > 0.0f) d = cross(1.0f, e12);
            else d = cross(e12, 1.0f);
        }
        if (dot(d, d) < FLT_EPSILON * FLT_EPSILON) break;
        SVertex& vertex = s.v[s.count];
        vertex.indexA = fixSupport(c, fA, mulT(xfA.q, -d));
        vertex.wA = mul(xfA, fixVert(c, fA, vertex.indexA));
        vertex.indexB = fixSupport(c, fB, mulT(xfB.q, d));
        vertex.wB = mul(xfB, fixVert(c, fB, vertex.indexB));
        vertex.w = vertex.wB - vertex.wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i)
            if (vertex.indexA == saveA[i] && vertex.indexB == saveB[i]) { duplicate = true; break; }
        if (duplicate) break;
        ++s.count;
    }
    // GetWitnessPoints + distance
    V2 pA, pB;
    if (s.count == 1) {
        pA = s.v[0].wA; pB = s.v[0].wB;
    } else if (s.count == 2) {
        pA = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA;
        pB = s.v[0].a * s.v[0].wB + s.v[1].a * s.v[1].wB;
    } else {
        pA = s.v[0].a * s.v[0].wA + s.v[1].a * s.v[1].wA + s.v[2].a * s.v[2].wA;
        pB = pA;
    }
    float distance = length(pA - pB);
    // WriteCache
    cache.metric = simplexMetric(s);
    cache.count = s.count;
    for (int i = 0; i < s.count; ++i) { cache.indexA[i] = s.v[i].indexA; cache.indexB[i] = s.v[i].indexB; }
    return distance;
}

// ---- b2TimeOfImpact ---------------------------------------------------------------------------------------------
struct SepFn {
    int type;  // 0 points, 1 faceA, 2 faceB
    V2 localPoint, axis;
};

DEV void sepInit(Ctx c, SepFn& f, const SimplexCache& cache, int fA, const Sweep& sA, int fB, const Sweep& sB, float t1) {
    XF xfA = sweepXF(sA, t1), xfB = sweepXF(sB, t1);
    if (cache.count == 1) {
        f.type = 0;
        V2 pointA = mul(xfA, fixVert(c, fA, cache.indexA[0]));
        V2 pointB = mul(xfB, fixVert(c, fB, cache.indexB[0]));
        f.axis = pointB - pointA;
        normalize(f.axis);
    } else if (cache.indexA[0] == cache.indexA[1]) {
        f.type = 2;
        V2 b1 = fixVert(c, fB, cache.indexB[0]);
        V2 b2 = fixVert(c, fB, cache.indexB[1]);
        f.axis = cross(b2 - b1, 1.0f);
        normalize(f.axis);
        V2 normal = mul(xfB.q, f.axis);
        f.localPoint = 0.5f * (b1 + b2);
        V2 pointB = mul(xfB, f.localPoint);
        V2 pointA = mul(xfA, fixVert(c, fA, cache.indexA[0]));
        float s = dot(pointA - pointB, normal);
        if (s < 0.0f) f.axis = -f.axis;
    } else {
        f.type = 1;
        V2 a1 = fixVert(c, fA, cache.indexA[0]);
        V2 a2 = fixVert(c, fA, cache.indexA[1]);
        f.axis = cross(a2 - a1, 1.0f);
        normalize(f.axis);
        V2 normal = mul(xfA.q, f.axis);
        f.localPoint = 0.5f * (a1 + a2);
        V2 pointA = mul(xfA, f.localPoint);
        V2 pointB = mul(xfB, fixVert(c, fB, cache.indexB[0]));
        float s = dot(pointB - pointA, normal);
        if (s < 0.0f) f.axis = -f.axis;
    }
}

// FindMinSeparation (find = true: choose support points) and Evaluate (find = false) share everything else
DEV float sepEval(Ctx c, const SepFn& f, int fA, const Sweep& sA, int fB, const Sweep& sB, int& indexA, int& indexB,
                  float t, bool find) {
    XF xfA = sweepXF(sA, t), xfB = sweepXF(sB, t);
    if (f.type == 0) {
        if (find) {
            indexA = fixSupport(c, fA, mulT(xfA.q, f.axis));
            indexB = fixSupport(c, fB, mulT(xfB.q, -f.axis));
        }
        V2 pointA = mul(xfA, fixVert(c, fA, indexA));
        V2 pointB = mul(xfB, fixVert(c, fB, indexB));
        return dot(pointB - pointA, f.axis);
    } else if (f.type == 1) {
        V2 normal = mul(xfA.q, f.axis);
        V2 pointA = mul(xfA, f.localPoint);
        if (find) {
            indexA = -1;
            indexB = fixSupport(c, fB, mulT(xfB.q, -normal));
        }
        V2 pointB = mul(xfB, fixVert(c, fB, indexB));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = mul(xfB.q, f.axis);
        V2 pointB = mul(xfB, f.localPoint);
        if (find) {
            indexB = -1;
            indexA = fixSupport(c, fA, mulT(xfA.q, -normal));
        }
        V2 pointA = mul(xfA, fixVert(c, fA, indexA));
        return dot(pointA - pointB, normal);
    }
}

// returns true when the output state is e_touching; t receives output.t
DEVN bool timeOfImpact(Ctx c, int fA, Sweep sweepA, int fB, Sweep sweepB, float& tOut) {
    const float tMax = 1.0f;
    tOut = tMax;
    sweepNormalize(sweepA);
    sweepNormalize(sweepB);
    const float totalRadius = B2G_POLYGON_RADIUS + B2G_POLYGON_RADIUS;
    const float target = fmaxT(B2G_LINEAR_SLOP, totalRadius - 3.0f * B2G_LINEAR_SLOP);
    const float tolerance = 0.25f * B2G_LINEAR_SLOP;
    float t1 = 0.0f;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    cache.metric = 0.0f;
    bool touching = false;
    for (;;) {
        XF xfA = sweepXF(sweepA, t1), xfB = sweepXF(sweepB, t1);
        float distance = distanceGJK(c, cache, fA, xfA, fB, xfB);
        if (distance <= 0.0f) { tOut = 0.0f; break; }                             // e_overlapped
        if (distance < target + tolerance) { touching = true; tOut = t1; break; }  // e_touching
        SepFn fcn;
        sepInit(c, fcn, cache, fA, sweepA, fB, sweepB, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = sepEval(c, fcn, fA, sweepA, fB, sweepB, indexA, indexB, t2, true);
            if (s2 > target + tolerance) { tOut = tMax; done = true; break; }  // e_separated
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = sepEval(c, fcn, fA, sweepA, fB, sweepB, indexA, indexB, t1, false);
            if (s1 < target - tolerance) { tOut = t1; done = true; break; }      // e_failed
            if (s1 <= target + tolerance) { touching = true; tOut = t1; done = true; break; }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIterCount;
                float s = sepEval(c, fcn, fA, sweepA, fB, sweepB, indexA, indexB, t, false);
                if (absT(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; }
                else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == 8) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { tOut = t1; break; }  // e_failed
    }
    return touching;
}

// ---- b2World::SolveTOI ----------------------------------------------------------------------------------------------
DEV void clearToiFlags(Ctx c, int body) {  // invalidate cached TOIs of every contact on `body`
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        if (mI(c, mfix(c, fA, MF_BODY)) == body || mI(c, mfix(c, fB, MF_BODY)) == body)
            stI(c, cslot(c, k, SC_FLAGS), ldI(c, cslot(c, k, SC_FLAGS)) & ~CF_TOI);
    }
}

DEVN void solveTOI(Ctx c) {
    const int nb = H().nb;
    // m_stepComplete is always true here (no sub-stepping mode): reset sweeps' alpha0 and the contacts' TOI cache
    bool candidates = false;
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int fl = ldI(c, cslot(c, k, SC_FLAGS));
        // clears toi flag and toiCount; m_toi = 1 is dead (the value is only read under the toi flag, which is set together
        // with a freshly computed m_toi)
        if (fl & ~(CF_TOUCHING | CF_ENABLED | CF_FILTER)) stI(c, cslot(c, k, SC_FLAGS), fl & (CF_TOUCHING | CF_ENABLED | CF_FILTER));
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        if (bodyType(c, mI(c, mfix(c, fA, MF_BODY))) != 2 || bodyType(c, mI(c, mfix(c, fB, MF_BODY))) != 2) candidates = true;
    }
    // sweeps' alpha0 = 0: island bodies got it with their c0 / a0 when the step began (solveIslands); the others (static,
    // sleeping) can only hold a non-zero alpha0 if a TOI loop ran since the last reset (bit1 of the world flags)
    const int wf = ldI(c, sslot(c, SS_FLAGS));
    if (wf & 2) {
        for (int b = 0; b < nb; ++b) stF(c, bslot(c, b, SB_ALPHA0), 0.0f);
        if (!candidates) stI(c, sslot(c, SS_FLAGS), wf & ~2);
    }
    if (!candidates) return;  // only contacts with a non-dynamic body are ever swept (no bullets in the reference)
    if ((wf & 2) == 0) stI(c, sslot(c, SS_FLAGS), wf | 2);

    for (;;) {
        int minContact = -1;
        float minAlpha = 1.0f;
        for (int k = contactCount(c) - 1; k >= 0; --k) {
            int fl = ldI(c, cslot(c, k, SC_FLAGS));
            if ((fl & CF_ENABLED) == 0) continue;
            if ((fl >> 8) > B2G_MAX_SUB_STEPS) continue;
            float alpha = 1.0f;
            if (fl & CF_TOI) {
                alpha = ldF(c, cslot(c, k, SC_TOI));
            } else {
                int bA, bB, fA, fB;
                contactBodies(c, k, bA, bB, fA, fB);
                int typeA = bodyType(c, bA), typeB = bodyType(c, bB);
                bool activeA = isAwake(c, bA) && typeA != 0;
                bool activeB = isAwake(c, bB) && typeB != 0;
                if (!activeA && !activeB) continue;
                bool collideA = typeA != 2;
                bool collideB = typeB != 2;
                if (!collideA && !collideB) continue;
                Sweep sA = ldSweep(c, bA), sB = ldSweep(c, bB);
                float alpha0 = sA.alpha0;
                if (sA.alpha0 < sB.alpha0) {
                    alpha0 = sB.alpha0;
                    sweepAdvance(sA, alpha0);
                    stSweep(c, bA, sA);
                } else if (sB.alpha0 < sA.alpha0) {
                    alpha0 = sA.alpha0;
                    sweepAdvance(sB, alpha0);
                    stSweep(c, bB, sB);
                }
                float beta;
                bool touching = timeOfImpact(c, fA, sA, fB, sB, beta);
                if (touching) alpha = fminT(alpha0 + (1.0f - alpha0) * beta, 1.0f);
                else alpha = 1.0f;
                stF(c, cslot(c, k, SC_TOI), alpha);
                stI(c, cslot(c, k, SC_FLAGS), fl | CF_TOI);
            }
            if (alpha < minAlpha) { minContact = k; minAlpha = alpha; }
        }
        if (minContact < 0 || 1.0f - 10.0f * FLT_EPSILON < minAlpha) break;

        int bA, bB, fA, fB;
        contactBodies(c, minContact, bA, bB, fA, fB);
        Sweep backup1 = ldSweep(c, bA), backup2 = ldSweep(c, bB);
        bodyAdvance(c, bA, minAlpha);
        bodyAdvance(c, bB, minAlpha);
#ifndef B2G_NO_HOOK_TOI
        if (updateContact(c, minContact) == 1 && PRM().recMode != 0) onBeginContact(c, minContact);
#else
        updateContact(c, minContact);
#endif
        {
            int fl = ldI(c, cslot(c, minContact, SC_FLAGS));
            fl &= ~CF_TOI;
            fl += 1 << 8;  // ++m_toiCount
            stI(c, cslot(c, minContact, SC_FLAGS), fl);
            if ((fl & CF_ENABLED) == 0 || (fl & CF_TOUCHING) == 0) {
                stI(c, cslot(c, minContact, SC_FLAGS), fl & ~CF_ENABLED);
                stSweep(c, bA, backup1);
                stSweep(c, bB, backup2);
                syncTransform(c, bA);
                syncTransform(c, bB);
                continue;
            }
        }
        wakeBody(c, bA);
        wakeBody(c, bB);

        // mini island: the two TOI bodies + their touching contacts with non-dynamic bodies
        unsigned char ibodies[kMaxBodies];
        unsigned char icontacts[B2G_MAX_TOI_CONTACTS];
        int nib = 0, nic = 0;
        uint64_t bodyFlag = 0ull;
        uint64_t contactFlag[2] = {0ull, 0ull};
        ibodies[nib++] = (unsigned char)bA;
        ibodies[nib++] = (unsigned char)bB;
        icontacts[nic++] = (unsigned char)minContact;
        bodyFlag |= (1ull << bA) | (1ull << bB);
        contactFlag[minContact >> 6] |= 1ull << (minContact & 63);
        for (int i = 0; i < 2; ++i) {
            int body = i == 0 ? bA : bB;
            if (bodyType(c, body) != 2) continue;
            for (int k = contactCount(c) - 1; k >= 0; --k) {
                int cA, cB, gA, gB;
                contactBodies(c, k, cA, cB, gA, gB);
                if (cA != body && cB != body) continue;
                if (nib == 2 * B2G_MAX_TOI_CONTACTS) break;
                if (nic == B2G_MAX_TOI_CONTACTS) break;
                if ((contactFlag[k >> 6] >> (k & 63)) & 1ull) continue;
                int other = cA == body ? cB : cA;
                if (bodyType(c, other) == 2) continue;  // only static / kinematic / bullet bodies join
                Sweep backup = ldSweep(c, other);
                if (((bodyFlag >> other) & 1ull) == 0ull) bodyAdvance(c, other, minAlpha);
#ifndef B2G_NO_HOOK_TOI
                if (updateContact(c, k) == 1 && PRM().recMode != 0) onBeginContact(c, k);
#else
                updateContact(c, k);
#endif
                int fl = ldI(c, cslot(c, k, SC_FLAGS));
                if ((fl & CF_ENABLED) == 0 || (fl & CF_TOUCHING) == 0) {
                    stSweep(c, other, backup);
                    syncTransform(c, other);
                    continue;
                }
                contactFlag[k >> 6] |= 1ull << (k & 63);
                icontacts[nic++] = (unsigned char)k;
                if ((bodyFlag >> other) & 1ull) continue;
                bodyFlag |= 1ull << other;
                if (bodyType(c, other) != 0) wakeBody(c, other);
                ibodies[nib++] = (unsigned char)other;
            }
        }

        // b2Island::SolveTOI with subStep.dt = (1 - minAlpha) * dt, 20 position iterations, no warm starting
        float h = (1.0f - minAlpha) * PRM().dt;
        for (int it = 0; it < 20; ++it) {
            float minSeparation = 0.0f;
            for (int i = 0; i < nic; ++i) minSeparation = contactSolvePos(c, icontacts[i], bA, bB, minSeparation);
            if (minSeparation >= -1.5f * B2G_LINEAR_SLOP) break;
        }
        stV2(c, bslot(c, bA, SB_C0X), ldV2(c, bslot(c, bA, SB_CX)));
        stF(c, bslot(c, bA, SB_A0), ldF(c, bslot(c, bA, SB_A)));
        stV2(c, bslot(c, bB, SB_C0X), ldV2(c, bslot(c, bB, SB_CX)));
        stF(c, bslot(c, bB, SB_A0), ldF(c, bslot(c, bB, SB_A)));
        for (int i = 0; i < nic; ++i) contactInit(c, icontacts[i], 1.0f, false);
        for (int it = 0, nv = PRM().velIters; it < nv; ++it)
            for (int i = 0; i < nic; ++i) contactSolveVel(c, icontacts[i]);
        for (int i = 0; i < nib; ++i) {
            integratePosition(c, ibodies[i], h, false);
            syncTransform(c, ibodies[i]);
        }

        uint64_t moved = 0ull;
        for (int i = 0; i < nib; ++i) {
            int body = ibodies[i];
            if (bodyType(c, body) != 2) continue;
            XF xf1;
            xf1.q = rotOf(ldF(c, bslot(c, body, SB_A0)));
            xf1.p = ldV2(c, bslot(c, body, SB_C0X)) - mul(xf1.q, localCenter(c, body));
            moved |= syncBodyFixtures(c, body, xf1, ldXF(c, body));
            clearToiFlags(c, body);
        }
        if (moved) stMoved(c, ldMoved(c) | moved);
        findNewContacts(c);
    }
}

}  // namespace b2g
